#!/bin/bash
# sub-cell digit of the NMS spatial sort: parity tests, then cfg4 with and without it, then the default bench step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py tests/test_gpu_cli.py -x -q -m gpu -k "nms or pipeline or chain or stage3 or cli" > gpurun_out/pytest_fine.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_fine.log
for fm in 16 8 32 1000000000; do
  echo "== merge cfg4, PG_NMS_FINE_MIN=$fm"
  PG_NMS_FINE_MIN=$fm timeout 300 python scripts/bench_merge_stress.py 2>/dev/null | cut -c1-330
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-corpus --no-e2e --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
