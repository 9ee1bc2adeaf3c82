#!/bin/bash
# round 2, GPU call H: ncu --set full with source for conv1 wgrad v3b (and conv2 wgrad for comparison)
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
python bench.py --no-cpu --no-graph --no-module --e2e-api engine --steps 3 --warmup 3 > gpurun_out/r2h_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'conv1_wgrad3_kernel|sw_wgrad_kernel' -s 12 -c 3 -o gpurun_out/r2h_wgrad python bench.py --no-cpu --no-graph --no-module --e2e-api engine --steps 3 --warmup 3 > gpurun_out/r2h_ncu.log 2>&1
tail -3 gpurun_out/r2h_ncu.log | cut -c1-300
ls -la gpurun_out/r2h_wgrad.ncu-rep
