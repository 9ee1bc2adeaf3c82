#!/bin/bash
# round 2, GPU call A2 (8 GPUs): replica check + weak-scaling bench at 8 and 4 ranks with the final exchange kernel
set -x
T=r2A
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/dp_check.py > gpurun_out/${T}_dp_check_8gpu.log 2>&1; tail -5 gpurun_out/${T}_dp_check_8gpu.log | cut -c1-300
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 200 --warmup 20 --no-cpu > gpurun_out/${T}_bench_${n}gpu.json 2> gpurun_out/${T}_bench_${n}gpu.err; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/${T}_bench_${n}gpu.json").read().strip().splitlines()[-1])
    print("N=$n", d["ms_per_step"], d["value"], d["e2e"], d["config"].get("replicas_identical"), d["breakdown_us"].get("reduce_then_adam_exchange"))
except Exception as e: print("parse failed", e)
P
tail -3 gpurun_out/${T}_bench_${n}gpu.err | cut -c1-300
done
