#!/bin/bash
# ncu --set full of the two heavy refit kernels (psb_lm_fit, topology (14,14), 64 x 64 grid); summaries made on the box
python tools/refit_one.py 14,14 64 3 > gpurun_out/plain_refit.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lm_normal -s 1 -c 1 -f -o gpurun_out/r02_lm_normal python tools/refit_one.py 14,14 64 3 > gpurun_out/ncu_lmn.log 2>&1
python tools/refit_one.py 14,14 64 3 > gpurun_out/plain_refit2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lm_solve -s 1 -c 1 -f -o gpurun_out/r02_lm_solve python tools/refit_one.py 14,14 64 3 > gpurun_out/ncu_lms.log 2>&1
for r in r02_lm_normal r02_lm_solve; do
  python tools/ncu_summary.py gpurun_out/$r.ncu-rep gpurun_out/$r > /dev/null
  rm -f gpurun_out/$r.ncu-rep
done
tail -n 2 gpurun_out/ncu_lmn.log gpurun_out/ncu_lms.log 2>/dev/null | tail -6; ls gpurun_out | grep r02_lm
