#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2t}
for dbg in 16 15; do
  BFL_TC_DEBUG=$dbg timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_dbg${dbg}.json 2> gpurun_out/${T}_dbg${dbg}.err
  python -c "
import json
d=json.load(open('gpurun_out/${T}_dbg${dbg}.json'))
print('debug', $dbg, 'ms/step %.1f' % d['ms_per_step'], d['roofline']['launch_ms'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
"
done
