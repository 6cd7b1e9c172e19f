#!/bin/bash
# GPU-box quick check: parity tests, then the bench line condensed to step / e2e / stage times.
python -m pytest tests -m gpu -x -q 2>&1 | tail -1
python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('step ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
for k,v in d['stages'].items(): print('  %-20s %.4f ms  frac %.3f' % (k, v['ms'], v['frac']))
"
