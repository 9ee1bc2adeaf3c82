#!/bin/bash
# round 2, GPU call B2: configs[3] (obs 12) on the tensor cores -- parity tests, then the stacked12 bench in both modes
set -x
T=${1:-r2B}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stacked12.py -q -m gpu --timeout 600 -rf -s > gpurun_out/${T}_pytest12.log 2>&1; tail -25 gpurun_out/${T}_pytest12.log | cut -c1-600
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_step.py tests/test_gpu_parity.py -q -m gpu --timeout 600 -rf -x > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log | cut -c1-400
for m in bf16 fp32; do
timeout 300 python bench.py --workload stacked12 --mode $m --steps 100 --warmup 5 > gpurun_out/${T}_bench_stacked12_$m.json 2> gpurun_out/${T}_stacked12_$m.err; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/${T}_bench_stacked12_$m.json").read().strip().splitlines()[-1])
    print("$m", d["ms_per_step"], d["value"], d["config"]["final_loss"], d["roofline"]["frac"])
except Exception as e: print("parse failed", e)
P
tail -3 gpurun_out/${T}_stacked12_$m.err | cut -c1-300
done
