#!/bin/bash
# round 2, GPU call C: full GPU suite on the compact conv1 gradient path, bench, ring-depth comparison of conv1's wgrad
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
BC_TEST_OUT=gpurun_out timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -rf -s > gpurun_out/r2c_pytest.log 2>&1; grep -n "bf16 step vs\|passed\|failed\|FAILED" gpurun_out/r2c_pytest.log | cut -c1-1800
timeout 600 python bench.py --steps 200 --no-cpu > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 1300 gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
BC_C1WG_RING=45 timeout 600 python bench.py --steps 100 --no-cpu --no-module --e2e-api engine > gpurun_out/r2c_bench_ring45.json 2> gpurun_out/r2c_bench_ring45.err; tail -c 700 gpurun_out/r2c_bench_ring45.json; tail -5 gpurun_out/r2c_bench_ring45.err
