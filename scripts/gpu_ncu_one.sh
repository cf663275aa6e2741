#!/bin/bash
# usage: gpu_ncu_one.sh <kernel-regex> [env assignments...]: one full ncu capture of a kernel inside the bench command
mkdir -p gpurun_out
k=$1; shift
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
env "$@" $CMD > gpurun_out/plain_$k.log 2>&1 &&
env "$@" ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -o gpurun_out/prof_$k -f $CMD > gpurun_out/ncu_$k.log 2>&1
tail -2 gpurun_out/ncu_$k.log
