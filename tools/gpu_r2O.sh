#!/bin/bash
# M = 64 MMAs in the shifted-window wgrad kernels: parity (the D layout in TMEM is an assumption to verify), bench
set -x
T=${1:-r2O}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -q -m gpu --timeout 600 -rf -k "wgrad or step or training" > gpurun_out/${T}_pytest.log 2>&1; tail -12 gpurun_out/${T}_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 300 --no-cpu --no-module --e2e-api engine > gpurun_out/${T}_bench.json 2>gpurun_out/${T}_bench.err; python - <<P
import json
d=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["breakdown_us"])
P
tail -3 gpurun_out/${T}_bench.err
