#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_jpeg.py -x -q -m gpu > gpurun_out/pytest_jpeg.log 2>&1; echo "jpeg tests rc=$?"
tail -3 gpurun_out/pytest_jpeg.log | cut -c1-250
timeout 300 python scripts/bench_jpeg.py 8 > gpurun_out/bench_jpeg.log 2>&1; echo "bench_jpeg rc=$?"
grep '"chunk_bytes": 256' gpurun_out/bench_jpeg.log | cut -c1-200
timeout 500 python bench.py --no-cpu-baseline --no-corpus --sustained-seconds 0 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('e2e pages/step', d['e2e']['pages_per_step'], round(d['e2e']['value']), d['e2e']['timeline_ms']['steps'])"
