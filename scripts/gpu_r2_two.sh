#!/bin/bash
# 2 GPUs: K6 through the C ABI, bench at N=2 (corpus hash must equal the N=1 hash), N=1 beside it
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scripts/check_hist_allreduce.py > gpurun_out/hist_allreduce_2gpu.json 2> gpurun_out/hist_allreduce_2gpu.err; echo "allreduce rc=$?"
cat gpurun_out/hist_allreduce_2gpu.json; tail -3 gpurun_out/hist_allreduce_2gpu.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/bench_n2.err
echo "(N=1 line: gpu_r2_final.sh)"
python - <<'P'
import json
for f in ('gpurun_out/bench_n2.json',):
    try:
        d=json.load(open(f))
        print(f, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'h2d GB/s/gpu', round(d['e2e']['h2d_gb_per_s_per_gpu'],1), 'corpus', d['corpus']['hist_sha256'], d['corpus']['exchange_ms'], 'sustained', d['roofline'].get('sustained',{}).get('frac'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
        print('  timeline', d['e2e'].get('timeline_ms',{}).get('steps'))
    except Exception as e:
        print(f, 'ERR', e)
P
