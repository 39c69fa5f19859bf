#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2f}
timeout 60 ./benchmarks/mma_probe > gpurun_out/${T}_mma_probe.txt 2>&1; echo "mma probe rc=$?"
cat gpurun_out/${T}_mma_probe.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:als_tc_kernel --launch-skip 1 -c 3 -f -o gpurun_out/${T}_tc \
   python bench.py --workload c2_small --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/${T}_ncu.stdout 2> gpurun_out/${T}_ncu.stderr; echo "ncu rc=$?"
