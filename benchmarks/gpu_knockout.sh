#!/bin/bash
# timing experiments: the tensor-core kernel with parts of its pipeline switched off (results are garbage by design)
mkdir -p gpurun_out
T=${1:-r2j}
for dbg in 0 1 2 4 8 3 5 6 7 15; do
  BFL_TC_DEBUG=$dbg timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_dbg${dbg}.json 2> gpurun_out/${T}_dbg${dbg}.err
  python -c "
import json
d=json.load(open('gpurun_out/${T}_dbg${dbg}.json'))
print('debug', $dbg, 'ms/step %.1f' % d['ms_per_step'], d['roofline']['launch_ms'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
"
done
