#!/bin/bash
# One gpurun call: GPU tests, headline bench (tensor-core default and SIMT-only), small diagnostics.
mkdir -p gpurun_out
T=${1:-r2a}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_bench_tc.json 2> gpurun_out/${T}_bench_tc.err; echo "bench tc rc=$?"
cat gpurun_out/${T}_bench_tc.json
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --kernel-mode 2 > gpurun_out/${T}_bench_simt.json 2> gpurun_out/${T}_bench_simt.err; echo "bench simt rc=$?"
cat gpurun_out/${T}_bench_simt.json
