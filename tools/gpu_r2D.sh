#!/bin/bash
# round 2, GPU call D2: obs-12 tests again (materialised batch through the second-generation wgrad), tensor-core suite, stacked12 at B = 256, 2-GPU stacked12
set -x
T=${1:-r2D}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stacked12.py tests/test_gpu_tc.py -q -m gpu --timeout 600 -rf > gpurun_out/${T}_pytest12.log 2>&1; tail -6 gpurun_out/${T}_pytest12.log | cut -c1-600
timeout 300 python bench.py --workload stacked12 --batch 255 --steps 100 --warmup 5 > gpurun_out/${T}_bench_stacked12_b255.json 2> gpurun_out/${T}_stacked12_b255.err; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/${T}_bench_stacked12_b255.json").read().strip().splitlines()[-1])
    print("B=255", d["ms_per_step"], d["value"], d["config"]["final_loss"], d["roofline"]["frac"])
except Exception as e: print("parse failed", e)
P
tail -3 gpurun_out/${T}_stacked12_b255.err | cut -c1-300
