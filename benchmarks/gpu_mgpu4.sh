#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2_mgpu4}
nvidia-smi -L | wc -l
timeout 400 python -m pytest tests/test_mgpu.py -m gpu -x -q -s > gpurun_out/${T}_test.log 2>&1; echo "mgpu test rc=$?"
grep "rank 0\|passed\|failed\|FAIL" gpurun_out/${T}_test.log | tail -12
