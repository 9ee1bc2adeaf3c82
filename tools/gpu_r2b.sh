#!/bin/bash
# round 2, GPU call B: re-run the failing tests, bench (module path), ncu --set full with source of the two conv1 kernels
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
BC_TEST_OUT=gpurun_out timeout 1200 python -m pytest tests/test_gpu_step.py tests/test_gpu_tc.py -q -m gpu --timeout 600 -rf -s > gpurun_out/r2b_pytest.log 2>&1; grep -n "bf16 step vs\|passed\|failed\|FAILED" gpurun_out/r2b_pytest.log | cut -c1-1500
timeout 600 python bench.py --steps 200 --no-cpu > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 900 gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
python bench.py --no-cpu --no-graph --no-module --e2e-api engine --steps 3 --warmup 3 > gpurun_out/r2b_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'conv1_tp_kernel|conv1_wgrad_tp_kernel' -s 8 -c 2 -o gpurun_out/r2b_conv1 python bench.py --no-cpu --no-graph --no-module --e2e-api engine --steps 3 --warmup 3 > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_ncu.log | cut -c1-300
ls -la gpurun_out/r2b_conv1.ncu-rep
