#!/bin/bash
# usage: gpu_cfg5.sh N   -> BASELINE configs[4] at N GPUs (and at 1 GPU for the hash comparison when N>1)
N=${1:-1}
mkdir -p gpurun_out
run() {
  n=$1
  if [ "$n" = "1" ]; then L="python bench.py"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py"; fi
  timeout 1200 $L --gpus $n --workload cfg5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/cfg5_n$n.log 2>&1
  echo "cfg5 N=$n exit $?"
  grep '^{"metric"' gpurun_out/cfg5_n$n.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  n_gpus %d steps %d pages/s %.0f ms/step %.3f corpus %s' % (d['n_gpus'], d['steps'], d['value'], d['ms_per_step'], d['corpus']))" || tail -15 gpurun_out/cfg5_n$n.log
}
run $N
if [ "$N" != "1" ]; then CUDA_VISIBLE_DEVICES=0 run 1; fi
