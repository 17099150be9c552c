#!/bin/bash
timeout 300 python bench.py --no-cpu-baseline --no-corpus --sustained-seconds 0 --steps 30 --e2e-trace-all 2> gpurun_out/e2e_trace.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); e=d['e2e']
print('e2e', round(e['value']), 'host ms/submit', round(e['host_ms_per_submit'],2), 'clocks', e['clocks']['sm_mhz'], e['clocks']['reasons'])"
python - <<'P'
import json
for l in open('gpurun_out/e2e_trace.err'):
    if l.startswith('{"e2e_trace_all"'):
        tl=json.loads(l)['e2e_trace_all']
        for i,(a,b,c,d) in enumerate(tl): print(i, 'copy %.1f-%.1f (%.1f)  compute %.1f-%.1f (%.1f)'%(a,b,b-a,c,d,d-c))
P
