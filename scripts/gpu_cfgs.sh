#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/bench_merge_stress.py 8 100000 > gpurun_out/cfg4.json 2> gpurun_out/cfg4.err; echo "cfg4 exit $?"; cat gpurun_out/cfg4.json; tail -3 gpurun_out/cfg4.err
timeout 900 python bench.py --workload cfg2 --pages-per-gpu 19 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/cfg2.json 2> gpurun_out/cfg2.err; echo "cfg2 exit $?"
tail -1 gpurun_out/cfg2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('cfg2 pages/s %.0f ms/step %.3f tiler frac %.3f iso %s e2e %s' % (d['value'], d['ms_per_step'], r['frac'], r.get('isolated',{}).get('frac'), d.get('e2e',{}).get('value')))" || tail -5 gpurun_out/cfg2.err
timeout 900 python bench.py --workload cfg2 --pages-per-gpu 76 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/cfg2b.json 2> gpurun_out/cfg2.err; echo "cfg2 (76 pages) exit $?"
tail -1 gpurun_out/cfg2b.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('cfg2x4 pages/s %.0f ms/step %.3f tiler frac %.3f iso %s' % (d['value'], d['ms_per_step'], r['frac'], r.get('isolated',{}).get('frac')))" || tail -5 gpurun_out/cfg2.err
