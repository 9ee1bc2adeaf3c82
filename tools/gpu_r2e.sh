#!/bin/bash
# round 2, GPU call E: third-generation conv1 wgrad: parity + timing vs the second generation
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
BC_TEST_OUT=gpurun_out timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_step.py -q -m gpu --timeout 300 -rf > gpurun_out/r2e_pytest.log 2>&1; tail -12 gpurun_out/r2e_pytest.log | cut -c1-600
timeout 600 python bench.py --steps 100 --no-cpu --no-module --e2e-api engine > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 700 gpurun_out/r2e_bench.json; tail -5 gpurun_out/r2e_bench.err
BC_C1WG_GEN=2 BC_C1WG_RING=45 timeout 600 python bench.py --steps 100 --no-cpu --no-module --e2e-api engine > gpurun_out/r2e_bench_gen2.json 2> gpurun_out/r2e_bench_gen2.err; tail -c 500 gpurun_out/r2e_bench_gen2.json; tail -5 gpurun_out/r2e_bench_gen2.err
