#!/bin/bash
set -x
T=r2Z2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_parity.py -q -m gpu --timeout 600 -rf -k "two_rank or peer or data_parallel" > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/dp_check.py > gpurun_out/${T}_dp_check_2gpu.log 2>&1; tail -3 gpurun_out/${T}_dp_check_2gpu.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 300 --no-cpu > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err
timeout 300 python bench.py --steps 300 --no-cpu --no-module --e2e-api engine > gpurun_out/${T}_bench_1gpu.json 2>/dev/null
python - <<P
import json
    try:
        d=json.loads(open("gpurun_out/${T}_"+f+".json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d.get("e2e",{}).get("value"), d["config"].get("replicas_identical"))
    except Exception as e: print(f, "parse failed", e)
P
