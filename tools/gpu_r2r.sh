#!/bin/bash
# round 2, GPU call R: swapped-role conv1 forward with 3 epilogue groups (16 warps) -- parity, bench
set -x
T=${1:-r2r}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_step.py -q -m gpu --timeout 600 -rf -x > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log | cut -c1-400
timeout 300 python bench.py --steps 200 --no-cpu --no-module --e2e-api engine > gpurun_out/${T}_bench_gen4.json 2>gpurun_out/${T}_bench_gen4.err; python - <<P
import json
d=json.loads(open("gpurun_out/${T}_bench_gen4.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["breakdown_us"])
P
tail -3 gpurun_out/${T}_bench_gen4.err
