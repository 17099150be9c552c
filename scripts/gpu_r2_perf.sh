#!/bin/bash
# tiler consumer-warp experiment, 30-tile set, box-kernel launch list, cfg4
mkdir -p gpurun_out
for cw in 8 4; do
  echo "== PG_TILER_CW=$cw tiler-only cfg3"
  PG_TILER_CW=$cw timeout 200 python bench.py --tiler-only --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-corpus --sustained-seconds 2 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); r=d['roofline']
print('ms/step',round(d['ms_per_step'],4),'frac',round(r['frac'],4),'sustained',round(r['sustained']['ms_per_step'],4),round(r['sustained']['frac'] or 0,4),r['sustained']['clocks']['sm_mhz'])"
  echo "== PG_TILER_CW=$cw whole step cfg3"
  PG_TILER_CW=$cw timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-corpus --sustained-seconds 2 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); r=d['roofline']
print('ms/step',round(d['ms_per_step'],4),'value',round(d['value']),'frac',round(r['frac'],4),'sustained',round(r['sustained']['ms_per_step'],4),round(r['sustained']['frac'] or 0,4))"
  echo "== PG_TILER_CW=$cw grids"
  PG_TILER_CW=$cw timeout 200 python scripts/bench_tiler_grids.py 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['grids'], round(d['ms_per_launch'],4), round(d['frac_of_measured_hbm_peak'],4))"
done
echo "== parity with CW=4"
PG_TILER_CW=4 timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_jpeg.py -x -q -m gpu -k "tiler or tile" 2>&1 | tail -3
echo "== box kernel launch list (no overlap)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/step_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-overlap --no-corpus --sustained-seconds 0 > gpurun_out/ncu_step.log 2>&1; echo "ncu rc=$?"
python - <<'P'
import csv, collections, re
rows=[r for r in csv.reader(open('gpurun_out/step_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[1:]:
    v=float(r[vi].replace(',','')); u=r[ui]
    v = v/1e3 if u in ('ns','nsecond') else v
    agg.setdefault(re.sub(r'\(.*','',r[ki])[:50],[]).append(v)
for k,v in agg.items(): print('%-52s n=%3d  last=%.1f us'%(k,len(v),v[-1]))
P
echo "== cfg4"
timeout 300 python scripts/bench_merge_stress.py 2>&1 | tail -1
