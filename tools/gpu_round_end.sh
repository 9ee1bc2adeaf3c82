#!/bin/bash
# artefacts at the final commit of the round: smoke, full GPU suite, default bench line, inference sweep (both serving paths up to B = 64),
# launch list of the serving forward at B = 1 / 8
set -x
T=${1:-r2K}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
BC_TEST_OUT=gpurun_out timeout 900 python -m pytest tests -q -m gpu --timeout 300 -rf > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${T}_pytest_gpu.log | cut -c1-300
timeout 400 python bench.py > gpurun_out/${T}_bench_bf16path.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench_bf16path.json; tail -3 gpurun_out/${T}_bench.err
timeout 300 python bench.py --workload infer --steps 200 > gpurun_out/${T}_bench_infer_sweep.json 2> gpurun_out/${T}_infer.err; tail -c 300 gpurun_out/${T}_bench_infer_sweep.json; tail -3 gpurun_out/${T}_infer.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_infer_launches.csv python tools/infer_launches.py > gpurun_out/${T}_ncu.log 2>&1; tail -1 gpurun_out/${T}_ncu.log
