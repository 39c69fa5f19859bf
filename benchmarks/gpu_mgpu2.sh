#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2_mgpu2}
nvidia-smi -L
timeout 900 python -m pytest tests/test_mgpu.py -m gpu -x -q -s > gpurun_out/${T}_test.log 2>&1; echo "mgpu test rc=$?"
grep "rank 0\|passed\|failed\|FAIL" gpurun_out/${T}_test.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cat gpurun_out/${T}_bench.json | cut -c1-900
