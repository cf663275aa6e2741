#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gather_grid -s 4 -c 1 -o gpurun_out/prof_gather_grid -f $CMD > gpurun_out/ncu_gather.log 2>&1
tail -2 gpurun_out/ncu_gather.log
