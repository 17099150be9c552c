#!/bin/bash
# usage: gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
for extra in "" "--corpus-stats"; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline $extra > gpurun_out/bench_n$N.log 2>&1
echo "N=$N [$extra] exit $?"
grep '^{"metric"' gpurun_out/bench_n$N.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('  n_gpus %d pages/s %.0f  ms/step %.3f  tiler frac %.3f  e2e %.0f  clocks %s' % (d['n_gpus'], d['value'], d['ms_per_step'], r['frac'], d.get('e2e',{}).get('value',0), d['clocks']))" || tail -20 gpurun_out/bench_n$N.log
done
