#!/bin/bash
# compute-sanitizer memcheck (and a bounded racecheck) over one small bf16 + fp32 training step, then the insurance
# runs of the final state: smoke, GPU suite, default bench line
set -x
T=${1:-r2F}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 240 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_step.py > gpurun_out/${T}_sanitizer_memcheck.log 2>&1; tail -6 gpurun_out/${T}_sanitizer_memcheck.log | cut -c1-200
BC_TEST_OUT=gpurun_out timeout 600 python -m pytest tests -q -m gpu --timeout 300 -rf > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${T}_pytest_gpu.log | cut -c1-300
timeout 400 python bench.py > gpurun_out/${T}_bench_bf16path.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench_bf16path.json; tail -3 gpurun_out/${T}_bench.err
timeout 100 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_step.py > gpurun_out/${T}_sanitizer_racecheck.log 2>&1; tail -6 gpurun_out/${T}_sanitizer_racecheck.log | cut -c1-200
