#!/bin/bash
# final 8-GPU call: weak scaling at 8 / 4 ranks (and 1 on the same box) with the final kernels
set -x
T=r2Z8
mkdir -p gpurun_out
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2957$n bench.py --gpus $n --steps 200 --warmup 20 --no-cpu --no-module --e2e-api engine > gpurun_out/${T}_bench_${n}gpu.json 2> gpurun_out/${T}_bench_${n}gpu.err
done
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu --no-module --e2e-api engine > gpurun_out/${T}_bench_1gpu.json 2> /dev/null
python - <<P
import json
for n in (1,4,8):
    try:
        d=json.loads(open(f"gpurun_out/${T}_bench_{n}gpu.json").read().strip().splitlines()[-1])
        print("N=%d"%n, d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"].get("replicas_identical"), d.get("breakdown_us",{}).get("reduce_then_adam_exchange"))
    except Exception as e: print(n, "parse failed", e)
P
