#!/bin/bash
# stage-3 merge after the blocked emit network and the alive-mask rounds: parity, cfg4 timing, launch list, bench step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py tests/test_gpu_cli.py tests/test_stage1_chain.py -x -q -m gpu > gpurun_out/pytest_merge2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_merge2.log
timeout 300 python scripts/bench_merge_stress.py 2>/dev/null | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:'nms_' -s 18 -c 6 --csv --log-file gpurun_out/merge2_launches.csv python scripts/bench_merge_stress.py > /dev/null 2>&1
grep -v "^==" gpurun_out/merge2_launches.csv | cut -d, -f5,15- | cut -c1-160
timeout 600 python bench.py --steps 20 --warmup 5 --no-corpus --no-e2e --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
