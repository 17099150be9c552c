#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py tests/test_gpu_cli.py tests/test_stage1_chain.py -x -q -m gpu 2>&1 | tail -2
for i in 1 2 3; do timeout 300 python scripts/bench_merge_stress.py 2>/dev/null | cut -c1-120; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-corpus --no-e2e --no-cpu-baseline --sustained-seconds 0 2>&1 | tail -1 | cut -c1-200
