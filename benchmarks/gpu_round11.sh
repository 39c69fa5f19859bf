#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2n}
BFL_TC_DEBUG=15 timeout 600 ncu --set full --clock-control none --import-source on -k regex:als_tc_kernel --launch-skip 1 -c 2 -f -o gpurun_out/${T}_skel \
   python bench.py --workload c2_small --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/${T}_ncu.stdout 2> gpurun_out/${T}_ncu.stderr; echo "ncu rc=$?"
