#!/bin/bash
# Evidence run at the round's head commit, ordered by importance so that a short GPU budget still yields the
# top items.  Every step is skipped once DEADLINE seconds (default 540) have passed since the start.
# Output layout = scripts/gpu_profiles.sh (gpurun_out/prof/), plus merge_full.ncu-rep for the cfg4 merge kernels.
mkdir -p gpurun_out/prof
O=gpurun_out/prof
T0=$(date +%s)
DEADLINE=${DEADLINE:-540}
left() { echo $(( DEADLINE - ($(date +%s) - T0) )); }
step() { # min_seconds_needed, label, command...
  need=$1; label=$2; shift 2
  if [ "$(left)" -lt "$need" ]; then echo "skip $label ($(left) s left)"; return 1; fi
  s=$(date +%s); "$@"; rc=$?; echo "$label exit $rc ($(( $(date +%s) - s )) s)"; return $rc
}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/smi_before.csv 2>&1
# 1. the default bench line with a clock log beside it
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/clocks.csv 2>/dev/null &
SMI=$!
step 0 "bench default" bash -c "timeout 400 python bench.py > $O/bench_default.json 2> $O/bench_default.err"
kill $SMI
# 2. cfg4 merge kernels, full set (north_star: SM / L1 throughput of the merge); the plain run first
MCMD="python scripts/bench_merge_stress.py"
step 60 "merge plain" bash -c "timeout 200 $MCMD > $O/merge_plain.json 2> $O/merge_plain.err" && \
step 90 "merge ncu full" bash -c "timeout 400 ncu --set full --clock-control none --import-source on -k regex:nms_ -s 18 -c 6 -f -o $O/merge_full $MCMD > $O/ncu_merge.log 2>&1"
# 3. launch list of the bench command (serialised under ncu: shares, not absolutes)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-overlap"
step 90 "launch plain" bash -c "timeout 200 $CMD > $O/plain_launches.log 2>&1" && \
step 60 "launch list" bash -c "timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1"
# 4. the other bench lines
step 40 "bench no-overlap" bash -c "timeout 200 python bench.py --no-overlap --no-cpu-baseline --no-e2e > $O/bench_no_overlap.json 2>> $O/bench_default.err"
step 40 "bench tiler-only" bash -c "timeout 200 python bench.py --tiler-only --no-cpu-baseline > $O/bench_tiler_only.json 2>> $O/bench_default.err"
step 50 "bench reference" bash -c "timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2>> $O/bench_default.err"
# 5. the dominant kernel, full set, at the bench's 64 pages per launch
CMD2="python bench.py --tiler-only --steps 1 --warmup 3 --no-cpu-baseline"
[ -n "$SKIP_TILER_NCU" ] || step 100 "tiler ncu full" bash -c "timeout $(( $(left) > 60 ? $(left) : 60 )) ncu --set full --clock-control none --import-source on -k regex:tile_letterbox -s 3 -c 1 -f -o $O/tiler_full $CMD2 > $O/ncu_tiler.log 2>&1"
# 6. optionally the GPU tests a change did not touch yet (PYTEST_K = a -k expression), in the time that is left
if [ -n "$PYTEST_K" ] && [ "$(left)" -gt 40 ]; then
  timeout $(left) python -m pytest tests -m gpu -q -k "$PYTEST_K" --maxfail=5 -p no:cacheprovider > gpurun_out/pytest_rest.log 2>&1
  echo "pytest rest exit $?"; tail -4 gpurun_out/pytest_rest.log
fi
ls -la $O
echo "total $(( $(date +%s) - T0 )) s"
