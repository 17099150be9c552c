#!/bin/bash
# Merge kernels after a change: the NMS parity tests, then the cfg4 stress timing.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -k "nms or chained or pipeline" --maxfail=5 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_nms.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_nms.log
timeout 120 python scripts/bench_merge_stress.py > gpurun_out/merge_stress.json 2> gpurun_out/merge_stress.err
echo "stress exit $?"; cat gpurun_out/merge_stress.json; tail -3 gpurun_out/merge_stress.err
PG_NMS_CLUSTER_MIN_BOXES=0 timeout 120 python scripts/bench_merge_stress.py > gpurun_out/merge_stress_nocluster.json 2>> gpurun_out/merge_stress.err
echo "stress (one CTA per page) exit $?"; cat gpurun_out/merge_stress_nocluster.json
# per-kernel durations of one call (serialised, cold-cache)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nms_ -s 18 -c 6 --csv --log-file gpurun_out/merge_launches.csv python scripts/bench_merge_stress.py > gpurun_out/ncu_merge_launches.log 2>&1
echo "ncu exit $?"; grep -v "^==" gpurun_out/merge_launches.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    print(r['Kernel Name'][:40], r['Metric Value'], r['Metric Unit'], r.get('Grid Size'))"
