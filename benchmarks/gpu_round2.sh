#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2b}
timeout 120 ./benchmarks/gather_probe > gpurun_out/${T}_gather_probe.txt 2>&1; echo "probe rc=$?"
cat gpurun_out/${T}_gather_probe.txt
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 400 > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/${T}_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:als_tc_kernel -c 2 -f -o gpurun_out/${T}_tc \
   python bench.py --workload c2_small --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/${T}_ncu.stdout 2> gpurun_out/${T}_ncu.stderr; echo "ncu rc=$?"
ls -la gpurun_out | tail -5
