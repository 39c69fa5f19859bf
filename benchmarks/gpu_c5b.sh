#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2_c5b}
timeout 900 python -m pytest tests/test_als_gpu.py -m gpu -x -q --timeout 300 -k "256" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/${T}_tests.log
timeout 1200 python bench.py --workload c5 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/${T}_n1.json 2> gpurun_out/${T}_n1.err; echo "c5 n1 rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/${T}_n1.json'))
print('c5 N=1: %.4g nnz/s, %.0f ms/iteration' % (d['value'], d['ms_per_step']), d['roofline']['launch_ms'], 'frac %.3f' % d['roofline']['frac'])
"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_c5small.csv python bench.py --workload c5_small --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${T}_l.stdout 2> gpurun_out/${T}_l.stderr; echo "launch list rc=$?"
