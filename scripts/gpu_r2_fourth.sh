#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_jpeg.py tests/test_stage1_chain.py -x -q -m gpu > gpurun_out/pytest_new.log 2>&1; echo "new tests rc=$?"
tail -5 gpurun_out/pytest_new.log | cut -c1-250
timeout 600 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_jpeg.py --deselect tests/test_stage1_chain.py > gpurun_out/pytest_rest.log 2>&1; echo "rest rc=$?"
tail -5 gpurun_out/pytest_rest.log | cut -c1-250
timeout 300 python scripts/bench_jpeg.py 8 > gpurun_out/bench_jpeg.log 2>&1; echo "bench_jpeg rc=$?"
grep -v '"chunk_bytes": 2048' gpurun_out/bench_jpeg.log | tail -22 | cut -c1-330
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_default.err
python - <<'P'
import json
try:
    d=json.load(open('gpurun_out/bench_default.json'))
    for k in ('value','ms_per_step','e2e'):
        print(k, json.dumps(d.get(k))[:1800])
except Exception as e:
    print('no bench line', e)
P
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/jpeg_launches.csv python scripts/bench_jpeg.py 8 once > gpurun_out/ncu_jpeg.log 2>&1; echo "ncu rc=$?"
python - <<'P'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/jpeg_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[1:]:
    v=float(r[vi].replace(',','')); u=r[ui]
    v = v/1e3 if u in ('ns','nsecond') else v
    agg.setdefault(r[ki][:60],[]).append(v)
for k,v in agg.items(): print('%-62s n=%3d  last=%.1f us  sum=%.1f us'%(k,len(v),v[-1],sum(v)))
P
