#!/bin/bash
# bench.py at N GPUs exactly as the driver launches it (torchrun, one rank per GPU), then the reference arm the same way
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2963$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
python - <<P
import json
d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1]); e=d['e2e']
print('N=$N value', round(d['value']), 'e2e', round(e['value']), 'h2d GB/s/gpu', round(e['h2d_gb_per_s_per_gpu'],1), 'corpus', d['corpus']['hist_sha256'], round(d['corpus']['exchange_ms'],3))
P
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2964$N bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | grep '^{' | cut -c1-260
