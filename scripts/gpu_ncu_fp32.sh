#!/bin/bash
# ncu --set full capture of the fp32-mode GEMM (tgemm_pair_kernel) inside a 128^3 fp32 grid evaluation; the plain command runs first.
mkdir -p gpurun_out
CMD="python scripts/bench_configs.py cfg3"
$CMD > gpurun_out/plain_fp32.log 2>&1 || { tail -5 gpurun_out/plain_fp32.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'tgemm_pair_kernel' -s 60 -c 1 -o gpurun_out/r02_prof_tgemm -f $CMD > gpurun_out/ncu_tgemm.log 2>&1
tail -2 gpurun_out/ncu_tgemm.log; tail -2 gpurun_out/plain_fp32.log | cut -c1-200
