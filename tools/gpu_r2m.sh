#!/bin/bash
# round 2, GPU call M (2 GPUs): exchange kernel with one float4 per thread; exchange timing in the breakdown
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_step.py -q -m gpu --timeout 600 -rf -k "two_rank" > gpurun_out/r2m_pytest.log 2>&1; tail -3 gpurun_out/r2m_pytest.log | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --no-cpu --no-module --e2e-api engine > gpurun_out/r2m_bench_2gpu.json 2> gpurun_out/r2m_bench_2gpu.err; tail -c 700 gpurun_out/r2m_bench_2gpu.json; tail -4 gpurun_out/r2m_bench_2gpu.err
timeout 300 python bench.py --steps 200 --no-cpu --no-module --e2e-api engine > gpurun_out/r2m_bench_1gpu.json 2>/dev/null; tail -c 400 gpurun_out/r2m_bench_1gpu.json
