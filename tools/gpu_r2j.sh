#!/bin/bash
# round 2, GPU call J: conv1 wgrad v3c smem plans (4 planes + 7 gradients vs 5 + 5), full GPU suite
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
BC_C1WG_PLAN=55 timeout 300 python -m pytest tests/test_gpu_tc.py -q -m gpu --timeout 300 -k "compact or whole or tail" > gpurun_out/r2j_pytest55.log 2>&1; tail -2 gpurun_out/r2j_pytest55.log
BC_TEST_OUT=gpurun_out timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -rf > gpurun_out/r2j_pytest.log 2>&1; tail -4 gpurun_out/r2j_pytest.log | cut -c1-600
for p in 47 55 47 55; do BC_C1WG_PLAN=$p python tools/c1wg_ablate.py 2>&1 | grep conv1_wgrad | sed "s/^/plan $p /"; done | tee gpurun_out/r2j_plans.txt
