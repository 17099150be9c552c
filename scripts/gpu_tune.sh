#!/bin/bash
mkdir -p gpurun_out
run() { # label, env..., -- args
  label=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --pages-per-gpu 64 --no-cpu-baseline --no-e2e $EXTRA > gpurun_out/tune.log 2>&1
  tail -1 gpurun_out/tune.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label: pages/s %.0f ms/step %.3f tiler_ms %.3f frac %.3f' % (d['value'], d['ms_per_step'], r['kernel_ms_per_launch'], r['frac']))" || tail -3 gpurun_out/tune.log
}
EXTRA=""
run "base s4 ipc8 boxfirst" A=1
run "s3" PG_TILER_STAGES=3
run "ipc4" PG_TILER_IPC=4
run "ipc16" PG_TILER_IPC=16
run "ipc2" PG_TILER_IPC=2
run "tilerfirst" PG_TILER_FIRST=1
run "s3 ipc4 tilerfirst" PG_TILER_STAGES=3 PG_TILER_IPC=4 PG_TILER_FIRST=1
EXTRA="--tiler-only"
run "tiler-only s4" A=1
run "tiler-only s3" PG_TILER_STAGES=3
run "tiler-only ipc16" PG_TILER_IPC=16
run "tiler-only ipc64" PG_TILER_IPC=64
