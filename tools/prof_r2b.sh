#!/bin/bash
# Round-2 (second half) evidence in one GPU call: GPU tests, bench lines, then ncu on the slot-ordered ABF kernel at
# 1e6 rows with HOOMD's slot -> tag array (class SC_TAGS).  ncu only after the same command exited 0 without it.
unset PSB_NO_GRAPH
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02b_pytest_gpu.txt
python bench.py > gpurun_out/r02b_bench_default.json 2> gpurun_out/bench_default.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench_steps20.json 2> gpurun_out/bench20.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_bench_reference.json 2> gpurun_out/bench_ref.err
export PSB_NO_GRAPH=1
python tools/run_case.py abf1m 12 > gpurun_out/plain_abf1m.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -f -o gpurun_out/r02b_tags1m python tools/run_case.py abf1m 12 > gpurun_out/ncu_tags.log 2>&1
python tools/run_case.py abf1m 12 > gpurun_out/plain_abf1m_2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k "regex:step_kernel" -s 6 -c 8 --csv --log-file gpurun_out/r02b_launches_tags1m.csv python tools/run_case.py abf1m 12 > gpurun_out/ncu_tags_l.log 2>&1
python tools/ncu_summary.py gpurun_out/r02b_tags1m.ncu-rep gpurun_out/r02b_tags1m > /dev/null
rm -f gpurun_out/r02b_tags1m.ncu-rep
cat gpurun_out/r02b_pytest_gpu.txt; tail -c 600 gpurun_out/r02b_bench_default.json; tail -n 2 gpurun_out/ncu_tags.log; ls -la gpurun_out | head -30
