#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/debug_columns.py > gpurun_out/debug_columns.log 2>&1; echo "debug exit $?"; cat gpurun_out/debug_columns.log | tail -45
timeout 600 python bench.py --steps 10 --warmup 3 --pages-per-gpu 64 --no-cpu-baseline --no-e2e > gpurun_out/bench64.log 2>&1; tail -1 gpurun_out/bench64.log
CMD="python bench.py --tiler-only --pages-per-gpu 16 --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_tiler.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tile_letterbox -s 3 -c 1 -f -o gpurun_out/tiler_full $CMD > gpurun_out/ncu_tiler.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/plain_tiler.log
