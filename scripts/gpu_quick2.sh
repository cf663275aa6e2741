#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "fused or compose or full_size" 2>&1 | tail -8 | tee gpurun_out/tests_fused.log
LIST_B200_FUSED_SKIP=0 bash scripts/gpu_fused_probe.sh 2>&1 | grep -E "skip=[0-9] quarter" 
