#!/bin/bash
# round 2, GPU call A: does the new library load (shared cudart), full GPU test suite, bench with/without the side-stream backward
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1
if ! grep -q "smoke ok" gpurun_out/r2a_smoke.log; then
  tail -5 gpurun_out/r2a_smoke.log
  echo "SHARED CUDART FAILED -> static"; rm -f carla_imitation_learning_b200/libbc_b200.so
  export BC_CUDART=static
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke_static.log 2>&1; tail -3 gpurun_out/r2a_smoke_static.log
fi
tail -2 gpurun_out/r2a_smoke.log
BC_TEST_OUT=gpurun_out timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -rf > gpurun_out/r2a_pytest.log 2>&1; tail -40 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 200 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
timeout 600 python bench.py --steps 200 --overlap 0 --no-cpu --e2e-api engine > gpurun_out/r2a_bench_serial.json 2> gpurun_out/r2a_bench_serial.err; tail -c 1200 gpurun_out/r2a_bench_serial.json; tail -5 gpurun_out/r2a_bench_serial.err
timeout 600 python bench.py --workload infer --steps 100 > gpurun_out/r2a_infer.json 2> gpurun_out/r2a_infer.err; tail -c 600 gpurun_out/r2a_infer.json; tail -3 gpurun_out/r2a_infer.err
