#!/bin/bash
# same-box A/B of two library builds: 128 against 256 head CTAs / partial copies (BC_HEAD_BLOCKS)
mkdir -p gpurun_out
for i in 1 2 3; do
  for v in h128 h256; do
    if [ $v = h128 ]; then export BC_LIB_PATH=$PWD/carla_imitation_learning_b200/libbc_b200_h128.so; else unset BC_LIB_PATH; fi
    timeout 120 python bench.py --no-cpu --no-module --e2e-api engine --steps 400 2>/dev/null | python -c "
import sys, json, ctypes, os
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
l = ctypes.CDLL(os.environ.get('BC_LIB_PATH') or 'carla_imitation_learning_b200/libbc_b200.so'); l.bc_partials_floats.restype = ctypes.c_size_t
print('$v run $i', 'partials', l.bc_partials_floats(4, 9), 'ms/step', round(d['ms_per_step'], 5), 'head', d['breakdown_us']['head_fwd_ce_bwd'], 'reduce', d['breakdown_us']['reduce_partials'])
" | tee -a gpurun_out/r2M_ab_head_blocks.log
  done
done
