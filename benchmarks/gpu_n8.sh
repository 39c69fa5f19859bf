#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2_n8}
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload c5 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/${T}_c5.json 2> gpurun_out/${T}_c5.err; echo "c5 n8 rc=$?"
tail -2 gpurun_out/${T}_c5.err | cut -c1-300
grep '^{' gpurun_out/${T}_c5.json | cut -c1-1200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --algo bpr --workload c3 --steps 3 --warmup 3 --no-cpu > gpurun_out/${T}_c3.json 2> gpurun_out/${T}_c3.err; echo "c3 n8 rc=$?"
tail -2 gpurun_out/${T}_c3.err | cut -c1-300
grep '^{' gpurun_out/${T}_c3.json | cut -c1-1500
