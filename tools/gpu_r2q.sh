#!/bin/bash
# round 2, GPU call Q: ncu full capture of the swapped-role conv1 forward
set -x
T=r2q
mkdir -p gpurun_out
python bench.py --no-cpu --no-graph --no-module --e2e-api engine --steps 3 --warmup 3 > gpurun_out/${T}_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv1_fwd4_kernel -s 4 -c 1 -o gpurun_out/${T}_c1f4 python bench.py --no-cpu --no-graph --no-module --e2e-api engine --steps 3 --warmup 3 > gpurun_out/${T}_ncu.log 2>&1; tail -2 gpurun_out/${T}_ncu.log | cut -c1-200
