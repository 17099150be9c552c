#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"
grep -v "^\*\|OMP_NUM" gpurun_out/bench_n8.err | tail -5
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/bench_n8.json') if l.startswith('{')][-1])
e=d['e2e']
print('N=8 value', round(d['value']), 'e2e', round(e['value']), 'h2d GB/s/gpu', round(e['h2d_gb_per_s_per_gpu'],1), 'corpus', d['corpus']['hist_sha256'], round(d['corpus']['exchange_ms'],3), 'host', e.get('host'))
print(' timeline', e['timeline_ms']['steps'])
P
