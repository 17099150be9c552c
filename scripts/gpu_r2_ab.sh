#!/bin/bash
# A/B at the head commit: (1) the whole GPU suite on the new build, (2) tiler with / without the predicated tail-word
# load (PG_TILER_TAIL_PRED, second build in _variants/), (3) cfg4 merge with 8- and 16-wide clusters.
mkdir -p gpurun_out
V=multimodal_embeddings_b200/_variants
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for lib in "" "$V/libpagegeom_tail0.so"; do
  echo "== tiler, lib=${lib:-head}"
  PAGEGEOM_LIB=$lib timeout 300 python scripts/bench_tiler_grids.py --pages 32 --steps 20 2>&1 | grep grids | cut -c1-260
  PAGEGEOM_LIB=$lib timeout 600 python bench.py --steps 20 --warmup 5 --no-corpus --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/ab_bench_${lib:+tail0}.json
  python - <<P
import json
d=json.load(open("gpurun_out/ab_bench_${lib:+tail0}.json"))
print({k:d.get(k) for k in ("value","ms_per_step")}, d["roofline"].get("frac"), d["roofline"].get("sustained"))
P
done
for cm in 16 8; do
  echo "== merge cfg4, PG_NMS_CLUSTER_MAX=$cm"
  PG_NMS_DEBUG=1 PG_NMS_CLUSTER_MAX=$cm timeout 300 python scripts/bench_merge_stress.py 2>&1 | grep -v "^$" | sort | uniq -c | cut -c1-220
done
