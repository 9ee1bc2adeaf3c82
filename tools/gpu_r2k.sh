#!/bin/bash
# round 2, GPU call K (2 GPUs): the data-parallel tests and bench lines (bucketed, double-buffered peer exchange)
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_step.py -q -m gpu --timeout 600 -rf -k "two_rank or module_path_data_parallel" > gpurun_out/r2k_pytest.log 2>&1; tail -15 gpurun_out/r2k_pytest.log | cut -c1-1200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > gpurun_out/r2k_dp_check.log 2>&1; tail -6 gpurun_out/r2k_dp_check.log | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --no-cpu > gpurun_out/r2k_bench_2gpu.json 2> gpurun_out/r2k_bench_2gpu.err; tail -c 1500 gpurun_out/r2k_bench_2gpu.json; tail -8 gpurun_out/r2k_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 200 --no-cpu --dp-overlap 0 --no-module --e2e-api engine > gpurun_out/r2k_bench_2gpu_serial.json 2> gpurun_out/r2k_bench_2gpu_serial.err; tail -c 600 gpurun_out/r2k_bench_2gpu_serial.json; tail -8 gpurun_out/r2k_bench_2gpu_serial.err
timeout 300 python bench.py --steps 200 --no-cpu --no-module --e2e-api engine > gpurun_out/r2k_bench_1gpu.json 2>/dev/null; tail -c 300 gpurun_out/r2k_bench_1gpu.json
