#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe): launch list + full capture of the two hot kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gather_fwd -s 4 -c 2 -o gpurun_out/prof_gather -f $CMD > gpurun_out/ncu_gather.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 4 -c 2 -o gpurun_out/prof_mlp -f $CMD > gpurun_out/ncu_mlp.log 2>&1
tail -3 gpurun_out/plain.log; tail -5 gpurun_out/ncu_gather.log; tail -5 gpurun_out/ncu_mlp.log; ls -la gpurun_out
