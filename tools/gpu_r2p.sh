#!/bin/bash
# round 2, GPU call P: swapped-role conv1 forward (conv1_fwd4.cu) -- parity suite, then A/B against the third generation
set -x
T=r2p
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_step.py tests/test_gpu_parity.py -q -m gpu --timeout 600 -rf -x > gpurun_out/${T}_pytest.log 2>&1; tail -15 gpurun_out/${T}_pytest.log | cut -c1-400
timeout 300 python bench.py --steps 200 --no-cpu --no-module --e2e-api engine > gpurun_out/${T}_bench_gen4.json 2>gpurun_out/${T}_bench_gen4.err; tail -c 1500 gpurun_out/${T}_bench_gen4.json; tail -3 gpurun_out/${T}_bench_gen4.err
BC_C1FW_GEN=3 timeout 300 python bench.py --steps 200 --no-cpu --no-module --e2e-api engine > gpurun_out/${T}_bench_gen3.json 2>/dev/null; tail -c 1500 gpurun_out/${T}_bench_gen3.json
