#!/bin/bash
# decoder after the sixteen-bytes-at-once unstuff tests and the leaner staging loop: parity, batch timing, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_jpeg.py tests/test_gpu_pipeline.py -x -q -m gpu > gpurun_out/pytest_jpeg4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_jpeg4.log
timeout 600 python scripts/bench_jpeg.py 8 2>&1 | grep -v "^$" | cut -c1-330
timeout 300 python scripts/bench_jpeg.py 8 once > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/jpeg4_launches.csv python scripts/bench_jpeg.py 8 once > /dev/null 2>&1
python - <<'P'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/jpeg4_launches.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split("(")[0]; agg.setdefault(name,[0,0.0]); agg[name][0]+=1; agg[name][1]+=float(r[-1])
for k,(n,t) in agg.items(): print(f"{k:50s} {n:4d} {t/1000 if t>1e4 else t:10.1f}")
P
