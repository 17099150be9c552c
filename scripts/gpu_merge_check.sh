#!/bin/bash
# Merge kernels after a change: the NMS parity tests, then the cfg4 stress timing.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -k "nms" --maxfail=5 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_nms.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_nms.log
timeout 120 python scripts/bench_merge_stress.py > gpurun_out/merge_stress.json 2> gpurun_out/merge_stress.err
echo "stress exit $?"; cat gpurun_out/merge_stress.json; tail -3 gpurun_out/merge_stress.err
PG_NMS_CLUSTER_MIN_BOXES=0 timeout 120 python scripts/bench_merge_stress.py > gpurun_out/merge_stress_nocluster.json 2>> gpurun_out/merge_stress.err
echo "stress (one CTA per page) exit $?"; cat gpurun_out/merge_stress_nocluster.json
