#!/bin/bash
mkdir -p gpurun_out
export LIST_B200_FUSED=1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fused" 2>&1 | tail -3
for skip in 7 0; do LIST_B200_FUSED_SKIP=$skip COUNT=4194304 python scripts/fused_small.py 2>&1 | tail -1; done
LIST_B200_FUSED_SKIP=7 python scripts/fused_small.py > gpurun_out/plain.log 2>&1 &&
LIST_B200_FUSED_SKIP=7 ncu --set full --clock-control none --import-source on -k regex:sdf_fused -s 1 -c 1 -o gpurun_out/prof_fused -f python scripts/fused_small.py > gpurun_out/ncu_fused.log 2>&1
tail -n 1 gpurun_out/ncu_fused.log
