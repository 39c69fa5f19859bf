#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2_c3}
timeout 600 python bench.py --algo bpr --workload c3 --steps 3 --warmup 3 --no-cpu > gpurun_out/${T}_n1.json 2> gpurun_out/${T}_n1.err; echo "c3 n1 rc=$?"
tail -2 gpurun_out/${T}_n1.err | cut -c1-300
grep '^{' gpurun_out/${T}_n1.json | cut -c1-1800
