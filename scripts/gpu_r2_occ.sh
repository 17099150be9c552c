#!/bin/bash
for o in 3 4; do echo "== PG_NMS_MASK_OCC=$o"; PG_NMS_MASK_OCC=$o timeout 300 python scripts/bench_merge_stress.py 2>/dev/null | cut -c1-140; done
