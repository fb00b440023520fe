#!/bin/bash
# sweep the plane-kernel launch parameters (planes per CTA, threads) on one workload
W=${1:-default}; shift
for cfg in "$@"; do
  set -- ${cfg//,/ }
  RSG_KC_PLANE=$1 RSG_PLANE_T=$2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $W 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']['per_kernel_ms']
print('$W KC=$1 T=$2', 'step %.4f'%d['ms_per_step'], 'plane %.4f col %.4f'%(r['k_plane_rp'], r['k_col_fused']))"
done
