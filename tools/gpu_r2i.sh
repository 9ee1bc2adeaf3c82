#!/bin/bash
# round 2, GPU call I: conv1 wgrad v3c (four issuers): parity, timing, ablation
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
BC_TEST_OUT=gpurun_out timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_step.py -q -m gpu --timeout 300 -rf > gpurun_out/r2i_pytest.log 2>&1; tail -12 gpurun_out/r2i_pytest.log | cut -c1-800
timeout 600 python bench.py --steps 100 --no-cpu --no-module --e2e-api engine > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -c 700 gpurun_out/r2i_bench.json; tail -5 gpurun_out/r2i_bench.err
rm -f carla_imitation_learning_b200/libbc_b200.so
BC_NVCC_EXTRA=-DBC_ABLATE python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
for a in 0 1 3 4 16 17 31; do BC_C1WG_ABLATE=$a python tools/c1wg_ablate.py 2>&1 | grep "conv1_wgrad"; done | tee gpurun_out/r2i_ablate.txt
rm -f carla_imitation_learning_b200/libbc_b200.so
