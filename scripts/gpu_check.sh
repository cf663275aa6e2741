#!/bin/bash
# Runs on the GPU box under gpurun: tests (tcgen05 kernels isolated in subprocesses first), then bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== tc tests" ; timeout 900 python -m pytest tests/test_gpu_mlp_tc.py -q -s 2>&1 | tail -40 | tee gpurun_out/tests_tc.log
echo "== parity tests" ; timeout 1500 python -m pytest tests/test_gpu_parity.py -q -s 2>&1 | tail -60 | tee gpurun_out/tests_parity.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench bf16" ; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench_bf16.log
echo "== bench fp32 128" ; timeout 900 python bench.py --steps 2 --warmup 3 --dtype fp32 --res 128 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_fp32.log
