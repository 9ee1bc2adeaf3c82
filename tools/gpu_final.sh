#!/bin/bash
# artefacts of the final state: smoke, full GPU suite, bench lines (train bf16 / fp32, infer, stacked12, reference arm),
# ncu launch list and one full capture of the step's kernels
set -x
T=${1:-r2Z}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
BC_TEST_OUT=gpurun_out timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -rf > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/${T}_bench_bf16path.json 2> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench_bf16path.json; tail -3 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${T}_bench_reference_cpu.json 2>/dev/null; tail -c 400 gpurun_out/${T}_bench_reference_cpu.json
timeout 600 python bench.py --workload infer --steps 200 > gpurun_out/${T}_bench_infer_sweep.json 2>/dev/null; tail -c 300 gpurun_out/${T}_bench_infer_sweep.json
timeout 600 python bench.py --workload stacked12 --steps 100 --warmup 5 > gpurun_out/${T}_bench_stacked12_bf16.json 2> gpurun_out/${T}_stacked12.err; tail -c 600 gpurun_out/${T}_bench_stacked12_bf16.json; tail -3 gpurun_out/${T}_stacked12.err
timeout 600 python bench.py --mode fp32 --steps 50 --no-cpu --no-module --e2e-api engine > gpurun_out/${T}_bench_fp32path.json 2>/dev/null; tail -c 300 gpurun_out/${T}_bench_fp32path.json
python bench.py --no-cpu --no-graph --no-module --e2e-api engine --steps 2 --warmup 3 > gpurun_out/${T}_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 75 -c 30 --csv --log-file gpurun_out/${T}_launches_bf16path.csv python bench.py --no-cpu --no-graph --no-module --e2e-api engine --steps 2 --warmup 3 > gpurun_out/${T}_ncu_ll.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:conv1_tp_kernel|conv1_wgrad3_kernel|stage_gray_tp_kernel|adam_tick_step_kernel|head_kernel|reduce_partials_kernel|sw_wgrad_kernel|sw_dgrad_kernel' -s 22 -c 11 -o gpurun_out/${T}_full python bench.py --no-cpu --no-graph --no-module --e2e-api engine --steps 3 --warmup 3 > gpurun_out/${T}_ncu_full.log 2>&1; tail -2 gpurun_out/${T}_ncu_full.log | cut -c1-200
timeout 240 ncu --graph-profiling graph --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/${T}_graph_level.csv python bench.py --steps 6 --warmup 3 --no-cpu --no-module --e2e-api engine > gpurun_out/${T}_ncu_graph.log 2>&1
grep '"graph"' gpurun_out/${T}_graph_level.csv | tail -4 | cut -c1-250
