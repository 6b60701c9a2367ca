#!/bin/bash
# ncu --set full captures of the other hot kernels: pair kernel (cfg3), slot-ordered step kernel (1e6 atoms),
# general step kernel with the TMA-staged hill sum (cfg4)
set -e
export PSB_NO_GRAPH=1
for c in cfg3 harmonic1m cfg4; do python tools/run_case.py $c 6 > gpurun_out/plain_$c.log 2>&1; done
ncu --set full --clock-control none --import-source on -k regex:coord_pairs -s 3 -c 1 -f -o gpurun_out/prof_pairs python tools/run_case.py cfg3 6 > gpurun_out/ncu_pairs.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 5 -c 1 -f -o gpurun_out/prof_slot python tools/run_case.py harmonic1m 8 > gpurun_out/ncu_slot.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 5 -c 1 -f -o gpurun_out/prof_cfg4 python tools/run_case.py cfg4 8 > gpurun_out/ncu_cfg4.log 2>&1
for f in pairs slot cfg4; do tail -n 2 gpurun_out/ncu_$f.log; done
