#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
for args in "--workload cfg2 --pages-per-gpu 19" "--workload cfg2 --pages-per-gpu 76" "--workload cfg2 --pages-per-gpu 19 --no-overlap" ""; do
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $args > gpurun_out/bench_b.log 2>&1
  tail -1 gpurun_out/bench_b.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('[$args] pages/s %.0f ms/step %.3f tiler frac %.3f iso %s e2e %s' % (d['value'], d['ms_per_step'], r['frac'], r.get('isolated',{}).get('frac'), d.get('e2e',{}).get('value')))" || tail -8 gpurun_out/bench_b.log
done
