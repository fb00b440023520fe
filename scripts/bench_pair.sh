#!/bin/bash
# default + x4 bench, compact report
for W in default x4; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload $W "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$W', 'step %.4f ms'%d['ms_per_step'], '%.1f Gcu/s'%(d['value']/1e9), 'e2e %.3f ms'%d['e2e']['ms_per_step'], {k:round(v,4) for k,v in r['per_kernel_ms'].items()}, r['kernel'], round(r['frac'],3))"
done
