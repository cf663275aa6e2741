#!/bin/bash
# ncu evidence for the line-table path (B200_PROFILING.md recipe), each pass only after the same command ran clean without ncu:
# launch list of the bench command, then ONE --set full capture of the four per-chunk kernels of the third chunk.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-library"
$CMD > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 400 --csv --log-file gpurun_out/r02_ncu_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'grid_tc_kernel|hoist_lines_kernel|hoist_rest_kernel|grid_plan_kernel' -s 8 -c 4 \
    -o gpurun_out/r02_prof_step -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
tail -1 gpurun_out/plain.log | cut -c1-300; ls -la gpurun_out/r02_prof_step.ncu-rep gpurun_out/r02_ncu_launches.csv
