#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
for mode in "" "--no-overlap" "--tiler-only"; do
  timeout 600 python bench.py --steps 10 --warmup 3 --pages-per-gpu 64 --no-cpu-baseline --no-e2e $mode > gpurun_out/bench_v2.log 2>&1
  echo "mode [$mode] exit $?"
  tail -1 gpurun_out/bench_v2.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('  pages/s %.0f  ms/step %.3f  tiler ms %.3f frac %.3f share %.2f' % (d['value'], d['ms_per_step'], r['kernel_ms_per_launch'], r['frac'], r['kernel_share_of_step']))" || tail -5 gpurun_out/bench_v2.log
done
