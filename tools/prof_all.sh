#!/bin/bash
# One GPU call: plain bench run, launch list (gpu__time_duration per launch) and a full ncu capture of the
# step kernel -> gpurun_out/.  Summaries for profiles/ are made afterwards with tools/profile_summaries.sh
set -e
CMD="python bench.py --steps 40 --warmup 5 --profile"
export PSB_NO_GRAPH=1
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 20 -c 2 -f -o gpurun_out/prof_step $CMD > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
