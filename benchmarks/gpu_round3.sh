#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2c}
timeout 60 ./benchmarks/mma_probe > gpurun_out/${T}_mma_probe.txt 2>&1; echo "mma probe rc=$?"
cat gpurun_out/${T}_mma_probe.txt
timeout 900 python -m pytest tests/test_als_gpu.py -m gpu -x -q --timeout 300 > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cat gpurun_out/${T}_bench.json
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --tc-min-class 0 > gpurun_out/${T}_bench_c0.json 2> gpurun_out/${T}_bench_c0.err; echo "bench rc=$?"
cat gpurun_out/${T}_bench_c0.json
