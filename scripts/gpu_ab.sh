#!/bin/bash
# A/B harness for a tiler change: parity tests first, then burst (20 steps) and sustained (1500 steps,
# power-capped clocks) timings of the tiler alone and of the whole step.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
run() { # label, args...
  label=$1; shift
  timeout 600 python bench.py --warmup 3 --pages-per-gpu 64 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ab.log 2>&1
  tail -1 gpurun_out/ab.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; c=d['clocks']; print('$label: pages/s %.0f ms/step %.3f tiler_ms %.3f frac %.3f iso %.3f sm_mhz %s %s' % (d['value'], d['ms_per_step'], r['kernel_ms_per_launch'], r['frac'], (r.get('isolated') or {}).get('frac', 0), c['sm_mhz'], c['reasons']))" || tail -3 gpurun_out/ab.log
}
run "tiler-only burst" --tiler-only --steps 20
run "tiler-only sustained" --tiler-only --steps 1500
run "step burst" --steps 20
run "step sustained" --steps 1500
