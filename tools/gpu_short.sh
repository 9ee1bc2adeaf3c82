#!/bin/bash
# short GPU call: full GPU suite + the default bench line (kernel-level changes show in breakdown_us)
set -x
T=${1:-r2s1}
mkdir -p gpurun_out
BC_TEST_OUT=gpurun_out timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -rf -x > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -8 gpurun_out/${T}_pytest_gpu.log | cut -c1-400
timeout 600 python bench.py --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d.get('module_api',{}).get('ms_per_step'))
print(d['breakdown_us'])
PY
