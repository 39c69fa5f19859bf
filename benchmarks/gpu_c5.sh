#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2_c5}
timeout 1200 python bench.py --workload c5 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/${T}_n1.json 2> gpurun_out/${T}_n1.err; echo "c5 n1 rc=$?"
tail -3 gpurun_out/${T}_n1.err
cat gpurun_out/${T}_n1.json | cut -c1-1500
