#!/bin/bash
# timing experiment: the Adam kernel with and without its two f64 pow (debug build on the box; results of the second build are wrong)
set -x
mkdir -p gpurun_out
run() { timeout 300 python bench.py --steps 300 --no-cpu --no-module --e2e-api engine > gpurun_out/r2K_$1.json 2>gpurun_out/r2K.err; python - <<P
import json
d=json.loads(open("gpurun_out/r2K_$1.json").read().strip().splitlines()[-1])
print("$1", d["ms_per_step"], d["breakdown_us"]["adam_tick_step_pack"])
P
}
run normal
touch carla_imitation_learning_b200/csrc/abi.cu
BC_NVCC_EXTRA=-DBC_ADAM_NOPOW python -c "from carla_imitation_learning_b200 import _lib; print(_lib.build())" | tail -1
run nopow
tail -2 gpurun_out/r2K.err
