#!/bin/bash
# ncu --set full at the head commit: decoder kernels, one-channel tiler, cfg4 merge kernels (each command ran without ncu first)
mkdir -p gpurun_out
timeout 300 python scripts/bench_jpeg.py 8 once > /dev/null 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:'jpeg_(store|spec|idct|unstuff_write|unstuff_count|colour)_kernel' -o gpurun_out/r02_jpeg_full -f python scripts/bench_jpeg.py 8 once > gpurun_out/ncu_full_jpeg.log 2>&1; echo "ncu jpeg rc=$?"
timeout 300 python scripts/bench_merge_stress.py > gpurun_out/merge_plain.json 2>/dev/null && \
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:'nms_' -s 18 -c 6 -o gpurun_out/r02_merge_full -f python scripts/bench_merge_stress.py > gpurun_out/ncu_full_merge.log 2>&1; echo "ncu merge rc=$?"
cat gpurun_out/merge_plain.json | cut -c1-300
ls -la gpurun_out/*.ncu-rep
