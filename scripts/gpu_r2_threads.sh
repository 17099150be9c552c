#!/bin/bash
# do narrower one-CTA-per-page box kernels disturb the tiler less?  (step time with the box stages on the other stream)
for t in 1024 512 256; do
  echo "== PG_BOX_PAGE_THREADS=$t"
  PG_BOX_PAGE_THREADS=$t timeout 600 python bench.py --steps 20 --warmup 5 --no-corpus --no-e2e --no-cpu-baseline --sustained-seconds 0 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'])"
done
PG_BOX_PAGE_THREADS=256 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -x -q -m gpu -k "nms or edge or pipeline or chained" 2>&1 | tail -2
