#!/bin/bash
# tests again + ncu launch list of a short bench (per-kernel device time; shares, not absolutes)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
CMD="python bench.py --steps 2 --warmup 3 --pages-per-gpu 64 --no-cpu-baseline --no-e2e"
timeout 900 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"
tail -2 gpurun_out/plain.log
