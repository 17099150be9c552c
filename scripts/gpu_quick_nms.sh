#!/bin/bash
# Shortest decisive check of a mask-kernel change: goldens, the synthetic batch at four thresholds (incl. thr < 0),
# the per-tile (class-agnostic fp32) mode, identical boxes; then the cfg4 timing.
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "nms_golden or nms_synthetic or per_tile_nms or nms_adversarial or nms_idempotent" > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_quick.log
timeout 20 python scripts/bench_merge_stress.py 2>&1 | tail -1 | cut -c1-200
