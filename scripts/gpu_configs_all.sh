#!/bin/bash
# Every BASELINE.json configuration plus the side benches, raw JSON lines into gpurun_out/configs/.
# usage: gpu_configs_all.sh [N]   (N > 1: the N-GPU legs only — cfg3 default + running exchange + cfg5)
N=${1:-1}
O=gpurun_out/configs
mkdir -p $O
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --no-cpu-baseline > $O/cfg3_n1.json 2> $O/err.log; echo "cfg3 exit $?"
  timeout 900 python bench.py --workload cfg2 --pages-per-gpu 19 --steps 10 --no-cpu-baseline > $O/cfg2_19pages.json 2>> $O/err.log; echo "cfg2 exit $?"
  timeout 900 python bench.py --workload cfg2 --pages-per-gpu 76 --steps 10 --no-cpu-baseline --no-e2e > $O/cfg2_76pages.json 2>> $O/err.log; echo "cfg2x4 exit $?"
  timeout 600 python scripts/bench_merge_stress.py 8 100000 > $O/cfg4_merge_stress.json 2>> $O/err.log; echo "cfg4 exit $?"
  timeout 900 python bench.py --workload cfg5 --total-pages 25600 --no-cpu-baseline --no-e2e > $O/cfg5_25600pages_n1.json 2>> $O/err.log; echo "cfg5 exit $?"
  timeout 900 python bench.py --records --no-cpu-baseline --no-e2e > $O/cfg3_records_n1.json 2>> $O/err.log; echo "records exit $?"
  timeout 900 python bench.py --graph --no-cpu-baseline --no-e2e > $O/cfg3_graph_n1.json 2>> $O/err.log; echo "graph exit $?"
  timeout 600 python scripts/bench_tiler_grids.py > $O/tiler_grid_sets.jsonl 2>> $O/err.log; echo "grids exit $?"
  timeout 600 python scripts/bench_json.py > $O/record_writer.json 2>> $O/err.log; echo "json exit $?"
  timeout 600 python scripts/bench_tile_nms.py > $O/tile_nms.json 2>> $O/err.log; echo "tile nms exit $?"
else
  L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
  timeout 900 $L bench.py --gpus $N --no-cpu-baseline > $O/cfg3_n$N.log 2>&1; echo "cfg3 N=$N exit $?"
  timeout 900 $L bench.py --gpus $N --no-cpu-baseline --no-e2e --corpus-stats > $O/cfg3_exchange_n$N.log 2>&1; echo "exchange N=$N exit $?"
  timeout 1200 $L bench.py --gpus $N --workload cfg5 --no-cpu-baseline --no-e2e > $O/cfg5_n$N.log 2>&1; echo "cfg5 N=$N exit $?"
  timeout 300 $L bench.py --gpus $N --impl reference --steps 1 --warmup 0 > $O/reference_arm_n$N.log 2>&1; echo "reference arm N=$N exit $?"
  for f in cfg3_n$N cfg3_exchange_n$N cfg5_n$N reference_arm_n$N; do grep -E '^\{"(metric|impl)"' $O/$f.log | tail -1 > $O/$f.json; done
fi
ls -la $O | tail -20
