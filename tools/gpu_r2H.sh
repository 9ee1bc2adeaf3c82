#!/bin/bash
# round 2, GPU call H2: "Toeplitz in N" dgrad of conv2 / conv3 -- tensor-core + step parity, bench
set -x
T=${1:-r2H}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_step.py tests/test_gpu_stacked12.py -q -m gpu --timeout 600 -rf > gpurun_out/${T}_pytest.log 2>&1; tail -8 gpurun_out/${T}_pytest.log | cut -c1-500
for i in 1 2; do
timeout 300 python bench.py --steps 400 --no-cpu --no-module --e2e-api engine > gpurun_out/${T}_bench$i.json 2>gpurun_out/${T}_bench.err; python - <<P
import json
d=json.loads(open("gpurun_out/${T}_bench$i.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["breakdown_us"])
P
done
tail -3 gpurun_out/${T}_bench.err
