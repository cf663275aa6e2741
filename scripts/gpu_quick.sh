#!/bin/bash
# quick iteration: parity tests + bench (bf16 256^3)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -25 | tee gpurun_out/tests_parity.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_bf16.log
LIST_B200_GRID_GENERIC=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_bf16_generic.log
