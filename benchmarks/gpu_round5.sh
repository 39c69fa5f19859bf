#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2e}
timeout 500 python -m pytest tests/test_als_gpu.py -m gpu -x -q --timeout 200 > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/${T}_tests.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cat gpurun_out/${T}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${T}_launches_c2.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_launches.stdout 2> gpurun_out/${T}_launches.stderr; echo "launch list rc=$?"
