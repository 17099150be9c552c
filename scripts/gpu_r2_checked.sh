#!/bin/bash
# the whole GPU suite + smoke against the checked build (-DPG_CHECKED: device-side bounds assertions, pg_common.cuh)
mkdir -p gpurun_out
L=$PWD/multimodal_embeddings_b200/_variants/libpagegeom_checked.so
ls -la $L
PAGEGEOM_LIB=$L timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_checked.log 2>&1; echo "pytest(checked) rc=$?"; tail -5 gpurun_out/pytest_checked.log
PAGEGEOM_LIB=$L timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
PAGEGEOM_LIB=$L timeout 300 python scripts/bench_merge_stress.py 2>/dev/null | cut -c1-200
