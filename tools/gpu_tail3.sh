#!/bin/bash
set -x
T=${1:-r2I}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 150 -k "policy_tail or inference_sweep" > gpurun_out/${T}_pytest_tail.log 2>&1; tail -4 gpurun_out/${T}_pytest_tail.log | cut -c1-300
timeout 300 python bench.py --workload infer --steps 200 > gpurun_out/${T}_bench_infer_sweep.json 2> gpurun_out/${T}_infer.err; python - <<P
import json
d=json.loads(open('gpurun_out/${T}_bench_infer_sweep.json').read().strip().splitlines()[-1])
for r in d['sweep'][:7]: print(r)
P
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_infer_launches.csv python tools/infer_launches.py > gpurun_out/${T}_ncu.log 2>&1; grep policy_tail gpurun_out/${T}_infer_launches.csv | head -3 | cut -c1-40,200-
