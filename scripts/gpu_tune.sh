#!/bin/bash
# Tiler tuning sweep (env knobs read by libpagegeom.so): ring depth, items per CTA, chunk threshold.
mkdir -p gpurun_out
run() { # label, env..., -- uses $EXTRA
  label=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --pages-per-gpu 64 --no-cpu-baseline --no-e2e $EXTRA > gpurun_out/tune.log 2>&1
  tail -1 gpurun_out/tune.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label: pages/s %.0f ms/step %.3f tiler_ms %.3f frac %.3f' % (d['value'], d['ms_per_step'], r['kernel_ms_per_launch'], r['frac']))" || tail -3 gpurun_out/tune.log
}
EXTRA="--tiler-only"
run "tiler-only default (s3 ipc4 chunk9216)" A=1
run "tiler-only s4" PG_TILER_STAGES=4
run "tiler-only ipc2" PG_TILER_IPC=2
run "tiler-only ipc8" PG_TILER_IPC=8
run "tiler-only chunk4700 (cfg3 tiles in two 512-px chunks)" PG_TILER_MAX_ROW_BYTES=4700
run "tiler-only chunk4700 s4" PG_TILER_MAX_ROW_BYTES=4700 PG_TILER_STAGES=4
EXTRA=""
run "step default" A=1
run "step ipc2" PG_TILER_IPC=2
run "step ipc8" PG_TILER_IPC=8
run "step chunk4700" PG_TILER_MAX_ROW_BYTES=4700
