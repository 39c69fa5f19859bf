#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2r}
timeout 500 python -m pytest tests/test_als_gpu.py -m gpu -x -q --timeout 200 > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/${T}_tests.log
for dbg in 0 4; do
  BFL_TC_DEBUG=$dbg timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_dbg${dbg}.json 2> gpurun_out/${T}_dbg${dbg}.err
  python -c "
import json
d=json.load(open('gpurun_out/${T}_dbg${dbg}.json'))
print('debug', $dbg, 'ms/step %.1f' % d['ms_per_step'], d['roofline']['launch_ms'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
"
done
