#!/bin/bash
mkdir -p gpurun_out
for ppg in 32 64 128 256; do
  timeout 900 python bench.py --steps 10 --warmup 3 --pages-per-gpu $ppg --no-cpu-baseline --no-e2e > gpurun_out/bench_ppg.log 2>&1
  tail -1 gpurun_out/bench_ppg.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('ppg $ppg: pages/s %.0f  ms/step %.3f  tiler ms %.3f frac %.3f' % (d['value'], d['ms_per_step'], r['kernel_ms_per_launch'], r['frac']))" || tail -5 gpurun_out/bench_ppg.log
done
