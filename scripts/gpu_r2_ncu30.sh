#!/bin/bash
# ncu --set full of the tiler on the reference's default grid set (30 tiles per page) and on cfg3's 4x4, head build
mkdir -p gpurun_out
timeout 300 python scripts/bench_tiler_grids.py --pages 32 --steps 3 --only default > /dev/null 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:tile_letterbox_kernel -s 3 -c 1 -o gpurun_out/r02_tiler30_full -f python scripts/bench_tiler_grids.py --pages 32 --steps 3 --only default > gpurun_out/ncu30.log 2>&1; echo "ncu30 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:tile_letterbox_kernel -s 3 -c 1 -o gpurun_out/r02_tiler44_full -f python scripts/bench_tiler_grids.py --pages 32 --steps 3 --only cfg3 > gpurun_out/ncu44.log 2>&1; echo "ncu44 rc=$?"
ls -la gpurun_out/r02_tiler*
