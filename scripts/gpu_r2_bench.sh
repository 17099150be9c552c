#!/bin/bash
mkdir -p gpurun_out
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/bench_default.json') if l.startswith('{')][-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'corpus',d['corpus']['hist_sha256'],'sust',d['roofline']['sustained']['frac'], 'dec', d['e2e']['jpeg_decoder'])
print(d['e2e']['timeline_ms']['steps'])
P
for pp in 8 32; do
timeout 300 python bench.py --no-cpu-baseline --no-corpus --sustained-seconds 0 --e2e-pages $pp 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('e2e pages/step', d['e2e']['pages_per_step'], round(d['e2e']['value']), d['e2e']['timeline_ms']['steps'])"
done
