#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2final2}
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print('value %.4g ms/step %.1f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']), d['roofline']['launch_ms'], 'e2e %.4g' % d['e2e']['value'], 'cpu %.4g x%d' % (d['cpu_baseline']['value'], d['cpu_baseline']['cores']), d['clocks'])
"
