#!/bin/bash
# the other BASELINE configurations at the round's last build, one GPU: cfg2 (19 page sizes in one batch), cfg5 (corpus streamed), cfg4 (merge stress)
mkdir -p gpurun_out
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --sustained-seconds 0 2>/dev/null | tail -1 > gpurun_out/r02_cfg2.json; echo "cfg2 rc=$?"
timeout 900 python bench.py --workload cfg5 --total-pages 20480 --no-cpu-baseline --no-e2e --sustained-seconds 0 2>/dev/null | tail -1 > gpurun_out/r02_cfg5.json; echo "cfg5 rc=$?"
timeout 300 python scripts/bench_merge_stress.py 2>/dev/null | tail -1 > gpurun_out/r02_cfg4.json; echo "cfg4 rc=$?"
python - <<'P'
import json
for n in ("cfg2","cfg5","cfg4"):
    try:
        d=json.loads(open(f"gpurun_out/r02_{n}.json").read())
        print(n, {k:d[k] for k in ("value","ms_per_step","ms_per_launch","pages_per_s") if k in d}, (d.get("config") or {}).get("workload","")[:80], (d.get("corpus") or {}).get("hist_sha256"))
    except Exception as e:
        print(n, "ERR", e)
P
