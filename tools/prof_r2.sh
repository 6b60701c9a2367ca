#!/bin/bash
# Round-2 ncu evidence, one GPU call (plain run of every command first; ncu only after it exited 0):
#   launch lists with --cache-control none (no flush between launches: per-launch durations comparable with
#   bench.py's value_plain regime), --set full captures for the byte / stall counters.
set -e
export PSB_NO_GRAPH=1
CMD="python bench.py --steps 40 --warmup 5 --profile"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k "regex:step_kernel" -s 10 -c 60 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 20 -c 1 -f -o gpurun_out/r02_step $CMD > gpurun_out/ncu_step.log 2>&1
python tools/run_case.py harmonic1m 12 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -f -o gpurun_out/r02_slot1m python tools/run_case.py harmonic1m 12 > gpurun_out/ncu_slot.log 2>&1
python tools/run_case.py harmonic1m 12 > gpurun_out/plain4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k "regex:step_kernel" -s 6 -c 8 --csv --log-file gpurun_out/r02_launches_slot1m.csv python tools/run_case.py harmonic1m 12 > gpurun_out/ncu_slot_l.log 2>&1
NWIN=32 python tools/batch_probe.py > gpurun_out/plain5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel_batch -s 20 -c 1 -f -o gpurun_out/r02_batch env NWIN=32 python tools/batch_probe.py > gpurun_out/ncu_batch.log 2>&1
NWIN=32 python tools/batch_probe.py > gpurun_out/plain6.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k "regex:step_kernel_batch" -s 20 -c 40 --csv --log-file gpurun_out/r02_launches_batch.csv env NWIN=32 python tools/batch_probe.py > gpurun_out/ncu_batch_l.log 2>&1
# summaries are made on the box (the reports are too large to bring back)
for r in r02_step r02_slot1m r02_batch; do
  python tools/ncu_summary.py gpurun_out/$r.ncu-rep gpurun_out/$r > /dev/null
  ncu -i gpurun_out/$r.ncu-rep --page details --csv > gpurun_out/${r}_details_all.csv 2>/dev/null || true
  rm -f gpurun_out/$r.ncu-rep
done
for f in gpurun_out/ncu_step.log gpurun_out/ncu_slot.log gpurun_out/ncu_batch.log; do tail -n 2 $f; done
ls -la gpurun_out/
