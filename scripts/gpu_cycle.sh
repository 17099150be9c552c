#!/bin/bash
# standard dev cycle: GPU tests, bench (overlap / single stream / tiler only), ncu launch list of the single-stream step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
for mode in "" "--no-overlap" "--tiler-only"; do
  timeout 600 python bench.py --steps 10 --warmup 3 --pages-per-gpu 64 --no-cpu-baseline --no-e2e $mode > gpurun_out/bench_cycle.log 2>&1
  tail -1 gpurun_out/bench_cycle.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('[$mode] pages/s %.0f  ms/step %.3f  tiler ms %.3f frac %.3f' % (d['value'], d['ms_per_step'], r['kernel_ms_per_launch'], r['frac']))" || tail -5 gpurun_out/bench_cycle.log
done
CMD="python bench.py --steps 2 --warmup 3 --pages-per-gpu 64 --no-cpu-baseline --no-e2e --no-overlap"
timeout 900 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"
