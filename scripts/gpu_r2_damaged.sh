#!/bin/bash
# damaged entropy data through the kernels (default and checked build), then the whole GPU suite once more
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_jpeg.py -x -q -m gpu -k "damaged" > gpurun_out/damaged.log 2>&1; tail -3 gpurun_out/damaged.log | cut -c1-220
PAGEGEOM_LIB=$PWD/multimodal_embeddings_b200/_variants/libpagegeom_checked.so timeout 600 python -m pytest tests/test_gpu_jpeg.py -x -q -m gpu -k "damaged" > gpurun_out/damaged_checked.log 2>&1; tail -3 gpurun_out/damaged_checked.log | cut -c1-220
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
