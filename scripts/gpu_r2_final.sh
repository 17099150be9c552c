#!/bin/bash
# round 2 evidence run (one B200): tests, smoke, bench lines, decoder timings, launch lists, command-line timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -3 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log | cut -c1-300
timeout 500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>&1; echo "ref rc=$?"
timeout 300 python scripts/bench_jpeg.py 8 > gpurun_out/bench_jpeg.log 2>&1; echo "bench_jpeg rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 200 --csv --log-file gpurun_out/jpeg_launches.csv python scripts/bench_jpeg.py 8 once > gpurun_out/ncu_jpeg.log 2>&1; echo "ncu jpeg rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/step_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-overlap --no-corpus --sustained-seconds 0 > gpurun_out/ncu_step.log 2>&1; echo "ncu step rc=$?"
timeout 400 python scripts/bench_cli_stage1.py --pages 16 2>/dev/null | tail -1 > gpurun_out/cli_stage1.json
for st in 2 3 4 5; do timeout 300 python scripts/bench_cli_stages.py --pages 8 --stage $st 2>/dev/null | tail -1; done > gpurun_out/cli_stages.jsonl
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/bench_default.json') if l.startswith('{')][-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'cpu',d['cpu_baseline']['value'],'corpus',d['corpus']['hist_sha256'],'sust',d['roofline']['sustained']['frac'])
P
