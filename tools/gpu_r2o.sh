#!/bin/bash
# round 2, GPU call O: full GPU suite after the stacked-camera window fix, configs[3] bench line, reference arm
set -x
T=r2o
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
BC_TEST_OUT=gpurun_out timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -rf > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --workload stacked12 --steps 50 --warmup 5 > gpurun_out/${T}_bench_stacked12.json 2> gpurun_out/${T}_stacked12.err; tail -c 900 gpurun_out/${T}_bench_stacked12.json; tail -3 gpurun_out/${T}_stacked12.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${T}_bench_reference_cpu.json 2>gpurun_out/${T}_ref.err; tail -c 500 gpurun_out/${T}_bench_reference_cpu.json; tail -3 gpurun_out/${T}_ref.err
