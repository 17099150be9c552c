#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
CMD="python bench.py --steps 1 --warmup 3 --pages-per-gpu 64 --no-cpu-baseline --no-e2e"
timeout 600 $CMD > gpurun_out/plain_nms.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:nms_pairs -s 7 -c 1 -f -o gpurun_out/nms_fill $CMD > gpurun_out/ncu_nms.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_nms.log
