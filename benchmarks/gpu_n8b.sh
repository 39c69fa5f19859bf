#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2_n8}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --algo bpr --workload c3 --steps 3 --warmup 3 --no-cpu > gpurun_out/${T}_c3.json 2> gpurun_out/${T}_c3.err; echo "c3 n8 rc=$?"
grep -i "error" gpurun_out/${T}_c3.err | head -3
grep '^{' gpurun_out/${T}_c3.json | cut -c1-1500
