#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_jpeg.py -x -q -m gpu > gpurun_out/pytest_jpeg.log 2>&1; echo "jpeg tests rc=$?"
tail -3 gpurun_out/pytest_jpeg.log | cut -c1-250
timeout 300 python scripts/bench_jpeg.py 8 > gpurun_out/bench_jpeg.log 2>&1; echo "bench_jpeg rc=$?"
grep -v '"chunk_bytes": 2048' gpurun_out/bench_jpeg.log | tail -22 | cut -c1-260
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 200 --csv --log-file gpurun_out/jpeg_launches.csv python scripts/bench_jpeg.py 8 once > gpurun_out/ncu_jpeg.log 2>&1; echo "ncu rc=$?"
python - <<'P'
import csv, collections, re
rows=[r for r in csv.reader(open('gpurun_out/jpeg_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit'); ii=hdr.index('ID')
agg=collections.OrderedDict()
for r in rows[1:]:
    k=re.sub(r'\(.*','',r[ki]).replace('<unnamed>::','')
    if not k.startswith('jpeg'): continue
    v=float(r[vi].replace(',',''))
    if r[mi]=='gpu__time_duration.sum' and r[ui] in ('ns','nsecond'): v/=1e3
    agg.setdefault((r[ii],k),{})[r[mi]]=v
for (i,k),m in agg.items():
    print('%-28s %8.1f us  lanes/inst %5.1f  issue %5.1f%%  warps %5.1f%%'%(k,m.get('gpu__time_duration.sum',0),m.get('smsp__thread_inst_executed_per_inst_executed.ratio',0),m.get('smsp__issue_active.avg.pct_of_peak_sustained_active',0),m.get('sm__warps_active.avg.pct_of_peak_sustained_active',0)))
P
