#!/bin/bash
# round 2, GPU call L: extension tests (branched heads, augmentation staging, folded BN)
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_extensions.py -q -m gpu --timeout 300 -rf > gpurun_out/r2l_pytest.log 2>&1; tail -40 gpurun_out/r2l_pytest.log | cut -c1-900
