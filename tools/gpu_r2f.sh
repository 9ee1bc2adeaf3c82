#!/bin/bash
# round 2, GPU call F: ablation of the third-generation conv1 wgrad
set -x
mkdir -p gpurun_out
rm -f carla_imitation_learning_b200/libbc_b200.so
BC_NVCC_EXTRA=-DBC_ABLATE python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
for a in 0 1 2 3 4 8 12 13 15 31 16; do BC_C1WG_ABLATE=$a python tools/c1wg_ablate.py 2>&1 | grep "conv1_wgrad"; done | tee gpurun_out/r2f_ablate.txt
rm -f carla_imitation_learning_b200/libbc_b200.so
