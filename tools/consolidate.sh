#!/bin/bash
# one GPU call: tests, bench (both arms), launch list and the full capture of the roofline kernel
set -x
BC_TEST_OUT=gpurun_out python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; tail -c 600 gpurun_out/bench_r1e.json
python bench.py --impl reference --steps 30 --warmup 3 > gpurun_out/bench_r1g_ref.json 2>/dev/null; tail -c 300 gpurun_out/bench_r1e_ref.json
python bench.py --no-cpu --no-graph --steps 2 --warmup 3 > gpurun_out/plain_ll.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/r1g_launches_bf16path.csv python bench.py --no-cpu --no-graph --steps 2 --warmup 3 > gpurun_out/ncu_ll.log 2>&1
python bench.py --no-cpu --no-graph --steps 3 --warmup 3 > gpurun_out/plain_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'conv1_tp_kernel|conv1_wgrad_tp_kernel|stage_gray_tp_kernel|sw_wgrad_kernel|sw_dgrad_kernel' -s 21 -c 7 -o gpurun_out/prof_r1g python bench.py --no-cpu --no-graph --steps 3 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
