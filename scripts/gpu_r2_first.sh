#!/bin/bash
# round 2, first GPU call: new tests, decoder timing, launch list of one decode
mkdir -p gpurun_out
python -m pytest tests/test_gpu_jpeg.py tests/test_stage1_chain.py -x -q -m gpu > gpurun_out/pytest_new.log 2>&1; echo "new tests rc=$?" 
tail -5 gpurun_out/pytest_new.log
python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_jpeg.py --deselect tests/test_stage1_chain.py > gpurun_out/pytest_rest.log 2>&1; echo "rest rc=$?"
tail -5 gpurun_out/pytest_rest.log
python scripts/bench_jpeg.py 8 > gpurun_out/bench_jpeg.log 2>&1; echo "bench_jpeg rc=$?"
cat gpurun_out/bench_jpeg.log | tail -30
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/bench_default.json
