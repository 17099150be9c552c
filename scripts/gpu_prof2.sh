#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --pages-per-gpu 64 --no-cpu-baseline --no-e2e --no-overlap"
timeout 600 $CMD > gpurun_out/plain_p2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"nms_mask|column_density" -s 6 -c 2 -f -o gpurun_out/box2 $CMD > gpurun_out/ncu_p2.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_p2.log
