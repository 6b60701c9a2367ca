#!/bin/bash
# launch list of this library's kernels during a short bench run (cold-cache, serialised: compare SHARES)
set -e
CMD="python bench.py --steps 40 --warmup 5 --profile"
export PSB_NO_GRAPH=1
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:step_kernel|coord_|inverse_permutation|grid_index|floor_kernel|latency_probe" -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
grep -c step_kernel gpurun_out/launches.csv
