#!/bin/bash
# soak: the same inputs through the kernels for tens of thousands of launches, every result compared with the first (scripts/soak.py)
mkdir -p gpurun_out
timeout 900 python scripts/soak.py --seconds ${1:-40} > gpurun_out/soak.jsonl 2> gpurun_out/soak.err; echo "soak rc=$?"; cat gpurun_out/soak.jsonl; tail -5 gpurun_out/soak.err
