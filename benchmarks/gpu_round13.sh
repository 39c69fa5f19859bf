#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2s}
timeout 500 python -m pytest tests/test_als_gpu.py -m gpu -x -q --timeout 200 > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/${T}_tests.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print('ms/step %.1f' % d['ms_per_step'], d['roofline']['launch_ms'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
"
