#!/bin/bash
# round 2, GPU call D: MMA micro-benchmark with mis-aligned A operands; conv1 wgrad with 128 B-aligned slices (both ring depths)
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
python tools/mma_bench.py > gpurun_out/r2d_mma_bench.txt 2>&1; tail -14 gpurun_out/r2d_mma_bench.txt
BC_TEST_OUT=gpurun_out timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_step.py -q -m gpu --timeout 600 -rf -x > gpurun_out/r2d_pytest.log 2>&1; tail -3 gpurun_out/r2d_pytest.log
timeout 600 python bench.py --steps 100 --no-cpu --no-module --e2e-api engine > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 700 gpurun_out/r2d_bench.json; tail -5 gpurun_out/r2d_bench.err
BC_C1WG_RING=45 timeout 600 python bench.py --steps 100 --no-cpu --no-module --e2e-api engine > gpurun_out/r2d_bench_ring45.json 2> gpurun_out/r2d_bench_ring45.err; tail -c 700 gpurun_out/r2d_bench_ring45.json; tail -5 gpurun_out/r2d_bench_ring45.err
