#!/bin/bash
# A/B of the mask kernel's occupancy knob (PG_NMS_MASK_OCC=3: 80 registers, 3 CTAs/SM; =4: 64 registers, 4 CTAs/SM).
mkdir -p gpurun_out
PG_NMS_MASK_OCC=4 timeout 100 python -m pytest tests -m gpu -q -k "nms" --maxfail=5 -p no:cacheprovider > gpurun_out/pytest_occ4.log 2>&1
echo "pytest (occ 4) exit $?"; tail -2 gpurun_out/pytest_occ4.log
for occ in 3 4; do
  PG_NMS_MASK_OCC=$occ timeout 60 python scripts/bench_merge_stress.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg4 occ $occ: %.4f ms' % d['ms_per_launch'])"
  PG_NMS_MASK_OCC=$occ timeout 60 python bench.py --no-overlap --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg3 one stream occ $occ: %.4f ms/step' % d['ms_per_step'])"
done
