#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2final3}
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${T}_launches_c2.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_launches.stdout 2> gpurun_out/${T}_launches.stderr; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:als_tc_kernel --launch-skip 1 -c 2 -f -o gpurun_out/${T}_tc \
   python bench.py --workload c2_small --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/${T}_ncu.stdout 2> gpurun_out/${T}_ncu.stderr; echo "ncu rc=$?"
