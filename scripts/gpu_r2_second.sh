#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_jpeg.py tests/test_stage1_chain.py -x -q -m gpu > gpurun_out/pytest_new.log 2>&1; echo "new tests rc=$?" 
tail -15 gpurun_out/pytest_new.log | cut -c1-250
python scripts/bench_jpeg.py 8 > gpurun_out/bench_jpeg.log 2>&1; echo "bench_jpeg rc=$?"
tail -30 gpurun_out/bench_jpeg.log | cut -c1-400
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_default.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_default.json'))
for k in ('value','ms_per_step','e2e','corpus','cpu_baseline'):
    print(k, json.dumps(d.get(k))[:900])
print('roofline', json.dumps({k:v for k,v in d['roofline'].items() if k in ('frac','isolated','sustained')})[:900])
P
