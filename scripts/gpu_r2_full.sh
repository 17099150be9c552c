#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -6 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log | cut -c1-300
timeout 500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/bench_default.json') if l.startswith('{')][-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'cpu',d['cpu_baseline']['value'],'corpus',d['corpus']['hist_sha256'],'sust',d['roofline']['sustained']['frac'], 'dec', d['e2e']['jpeg_decoder'])
P
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_reference.json | cut -c1-600
timeout 400 python scripts/bench_cli_stage1.py --pages 16 2>/dev/null | tail -1 > gpurun_out/cli_stage1.json; cat gpurun_out/cli_stage1.json
for st in 2 3 4 5; do timeout 300 python scripts/bench_cli_stages.py --pages 8 --stage $st 2>/dev/null | tail -1; done > gpurun_out/cli_stages.jsonl; cat gpurun_out/cli_stages.jsonl | cut -c1-400
