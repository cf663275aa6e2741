#!/bin/bash
mkdir -p gpurun_out
export LIST_B200_FUSED=1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "fused or compose or full_size" 2>&1 | tail -5 | tee gpurun_out/tests_fused.log
bash scripts/gpu_fused_probe.sh 2>&1 | grep -E "skip=[0-9] quarter"
