#!/bin/bash
# prints value, ms/step and the per-kernel split of one bench run
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-mpc --no-config4 --no-config5 --no-divergent "$@" 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']), round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['roofline']['ms_by_kernel_per_step'].items()}, 'e2e', round(d['e2e']['value']))
    else: print(l)
"
