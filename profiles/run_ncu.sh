#!/bin/bash
# Profiling recipe used for the numbers under profiles/ (run on the GPU box through gpurun, 1 GPU).
#   $1 = tag (e.g. r1_final)
# 1. launch list of the headline bench command with per-launch time and DRAM bytes (cold-cache, serialised:
#    compare SHARES, not absolutes)
# 2. one `--set full` capture of the dominant kernel family on the 1/10-scale workload (same kernels, 10x shorter)
set -x
TAG=${1:-rX}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches_c2.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_launches_c2.stdout 2> gpurun_out/${TAG}_launches_c2.stderr
ncu --set full --clock-control none --import-source on -k regex:als_tc_kernel --launch-skip 1 -c 2 -f -o gpurun_out/${TAG}_prof \
    python bench.py --workload c2_small --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/${TAG}_prof.stdout 2> gpurun_out/${TAG}_prof.stderr
ls -la gpurun_out/ | tail -5
