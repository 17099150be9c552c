#!/bin/bash
# Collect everything profiles/ needs, with the bench's own command lines.
mkdir -p gpurun_out/prof
O=gpurun_out/prof
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/smi_before.csv 2>&1
# 1. the default bench line (N=1) and the reference arm
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/clocks.csv 2>/dev/null &
SMI=$!
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench default exit $?"
kill $SMI
timeout 900 python bench.py --no-overlap --no-cpu-baseline > $O/bench_no_overlap.json 2>> $O/bench_default.err; echo "bench no-overlap exit $?"
timeout 900 python bench.py --tiler-only --no-cpu-baseline > $O/bench_tiler_only.json 2>> $O/bench_default.err; echo "bench tiler-only exit $?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>> $O/bench_default.err; echo "bench reference exit $?"
# 2. launch list of the same command (serialised under ncu: shares, not absolutes)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-overlap"
timeout 900 $CMD > $O/plain_launches.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1
echo "launch list exit $?"
# 3. the dominant kernel, full set, at the bench's 64 pages per launch
CMD2="python bench.py --tiler-only --steps 1 --warmup 3 --no-cpu-baseline"
timeout 900 $CMD2 > $O/plain_tiler.log 2>&1 && \
timeout 1800 ncu --set full --clock-control none --import-source on -k regex:tile_letterbox -s 3 -c 1 -f -o $O/tiler_full $CMD2 > $O/ncu_tiler.log 2>&1
echo "tiler full exit $?"
ls -la $O
