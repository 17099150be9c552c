#!/bin/bash
# e2e with the files staged in plain pinned memory vs write-combined pinned memory (N from the launcher)
mkdir -p gpurun_out
N=${1:-1}
for mode in default wc; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --steps 20 --warmup 3 --no-corpus --no-cpu-baseline --sustained-seconds 0 --e2e-pinned $mode 2>/dev/null | tail -1 > gpurun_out/wc_${mode}_n$N.json
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$N bench.py --gpus $N --steps 20 --warmup 3 --no-corpus --no-cpu-baseline --sustained-seconds 0 --e2e-pinned $mode 2>/dev/null | tail -1 > gpurun_out/wc_${mode}_n$N.json
  fi
  python - <<P
import json
d=json.loads(open("gpurun_out/wc_${mode}_n$N.json").read()); e=d["e2e"]
print("$mode N=$N e2e", round(e["value"]), "h2d/gpu", round(e["h2d_gb_per_s_per_gpu"],1), "host_ms", round(e["host_ms_per_submit"],2), e["timeline_ms"]["steps"][1])
P
done
