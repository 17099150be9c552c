#!/bin/bash
# First-contact GPU script: parity tests, smoke, short bench.  Everything under a timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 --pages-per-gpu 16 > gpurun_out/bench_small.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_small.log
tail -5 gpurun_out/bench_small.log
