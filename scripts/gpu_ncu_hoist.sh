#!/bin/bash
# ncu evidence for the hoisted path (B200_PROFILING.md recipe): launch list + full captures of the three hot kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --chunk 262144"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 400 --csv --log-file gpurun_out/launches_hoist.csv $CMD > gpurun_out/ncu_launches.log 2>&1
for k in hoist_addend hoist_rest mlp_tc; do
  $CMD > gpurun_out/plain_$k.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -o gpurun_out/prof_$k -f $CMD > gpurun_out/ncu_$k.log 2>&1
  tail -2 gpurun_out/ncu_$k.log
done
tail -1 gpurun_out/plain.log | cut -c1-400; ls -la gpurun_out | tail -8
