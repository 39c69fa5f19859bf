#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2_sgd}
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv \
    --log-file gpurun_out/${T}_c3_launches.csv python bench.py --algo bpr --workload c3 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${T}_c3.stdout 2> gpurun_out/${T}_c3.stderr; echo "c3 ncu rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv \
    --log-file gpurun_out/${T}_c4_launches.csv python bench.py --algo warp --workload c4 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${T}_c4.stdout 2> gpurun_out/${T}_c4.stderr; echo "c4 ncu rc=$?"
