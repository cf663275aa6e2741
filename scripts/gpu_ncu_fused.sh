#!/bin/bash
mkdir -p gpurun_out
export LIST_B200_FUSED=1
python scripts/fused_small.py > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sdf_fused -s 1 -c 1 -o gpurun_out/prof_fused -f python scripts/fused_small.py > gpurun_out/ncu_fused.log 2>&1
tail -n 2 gpurun_out/plain.log; tail -n 2 gpurun_out/ncu_fused.log
