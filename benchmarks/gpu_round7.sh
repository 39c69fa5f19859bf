#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2g}
timeout 500 python -m pytest tests/test_als_gpu.py -m gpu -x -q --timeout 200 > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/${T}_tests.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cat gpurun_out/${T}_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:als_tc_kernel --launch-skip 1 -c 2 -f -o gpurun_out/${T}_tc \
   python bench.py --workload c2_small --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/${T}_ncu.stdout 2> gpurun_out/${T}_ncu.stderr; echo "ncu rc=$?"
