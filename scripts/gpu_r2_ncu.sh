#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:'jpeg_(store|spec|idct|unstuff_write)_kernel' -o gpurun_out/r02_jpeg_full -f python scripts/bench_jpeg.py 8 once > gpurun_out/ncu_full_jpeg.log 2>&1; echo "ncu jpeg rc=$?"
tail -3 gpurun_out/ncu_full_jpeg.log
ls -la gpurun_out/*.ncu-rep
