#!/bin/bash
# the one-launch serving tail (csrc/policy_tail.cu): its tests, then the inference sweep with both paths at B <= 64
set -x
T=${1:-r2G}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 150 -k "policy_tail or inference_sweep or contract" > gpurun_out/${T}_pytest_tail.log 2>&1; tail -15 gpurun_out/${T}_pytest_tail.log | cut -c1-400
timeout 300 python bench.py --workload infer --steps 200 > gpurun_out/${T}_bench_infer_sweep.json 2> gpurun_out/${T}_infer.err; tail -c 2500 gpurun_out/${T}_bench_infer_sweep.json; tail -3 gpurun_out/${T}_infer.err
