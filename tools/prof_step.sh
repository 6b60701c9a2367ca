#!/bin/bash
# ncu capture of the fused step kernel in steady state (run under gpurun; writes gpurun_out/)
# usage: prof_step.sh [extra ncu args...]   e.g. --cache-control none
set -e
CMD="python bench.py --steps 40 --warmup 5 --profile"
export PSB_NO_GRAPH=1
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on "$@" -k regex:step_kernel -s 20 -c 2 -f -o gpurun_out/prof_step $CMD > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
