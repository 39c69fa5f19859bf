#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2v}
for c in 2 3; do
  timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --tc-min-class $c > gpurun_out/${T}_c${c}.json 2> gpurun_out/${T}_c${c}.err
  python -c "
import json
d=json.load(open('gpurun_out/${T}_c${c}.json'))
print('tc_min_class', $c, 'ms/step %.1f' % d['ms_per_step'], d['roofline']['launch_ms'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
"
done
