#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2h}
timeout 500 python -m pytest tests/test_als_gpu.py -m gpu -x -q --timeout 200 > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/${T}_tests.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_bench_n8.json 2> gpurun_out/${T}_bench_n8.err; echo "bench n8 rc=$?"
cat gpurun_out/${T}_bench_n8.json
cp buffalo_b200/libbuffalo_b200.so /tmp/n8.so; cp benchmarks/libbfl_n4.so buffalo_b200/libbuffalo_b200.so
timeout 400 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_bench_n4.json 2> gpurun_out/${T}_bench_n4.err; echo "bench n4 rc=$?"
cat gpurun_out/${T}_bench_n4.json
cp /tmp/n8.so buffalo_b200/libbuffalo_b200.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:als_tc_kernel --launch-skip 1 -c 2 -f -o gpurun_out/${T}_tc \
   python bench.py --workload c2_small --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/${T}_ncu.stdout 2> gpurun_out/${T}_ncu.stderr; echo "ncu rc=$?"
