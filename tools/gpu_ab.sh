#!/bin/bash
# same-box A/B of an environment switch on the graph-replayed step + whole-graph DRAM traffic (ncu --graph-profiling graph)
set -x
T=${1:-r2s2}; VAR=${2:-BC_STAGE_CS}
mkdir -p gpurun_out
for i in 1 2 3; do for f in 0 1; do
  env $VAR=$f timeout 300 python bench.py --steps 400 --no-cpu --no-module --e2e-api engine 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$VAR=$f', d['ms_per_step'], d['value'])"
done; done
for f in 0 1; do
env $VAR=$f timeout 240 ncu --graph-profiling graph --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/${T}_graph_$f.csv python bench.py --steps 6 --warmup 3 --no-cpu --no-module --e2e-api engine > gpurun_out/${T}_ncu_graph_$f.log 2>&1
grep -i "graph" gpurun_out/${T}_graph_$f.csv | tail -12 | cut -c1-300
done
