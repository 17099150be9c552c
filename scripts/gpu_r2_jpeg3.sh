#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_jpeg.py -x -q -m gpu > gpurun_out/pytest_jpeg.log 2>&1; echo "jpeg tests rc=$?"
tail -2 gpurun_out/pytest_jpeg.log | cut -c1-250
for t in smem global; do
echo "== tables $t"
PG_JPEG_TABLES=$t timeout 300 python scripts/bench_jpeg.py 8 2>&1 | grep '"chunk_bytes": 256' | cut -c1-170
done
