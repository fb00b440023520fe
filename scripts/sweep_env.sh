for cfg in "12 12 12 7" "6 6 6 7" "9 8 8 7" "6 12 12 5" "18 18 12 7" "12 12 12 35"; do
  set -- $cfg
  RSG_SEG_E=$1 RSG_SEG_MU=$2 RSG_SEG_P=$3 RSG_KC_R=$4 python bench.py --steps 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']['per_kernel_ms']
print('$cfg', 'step %.4f'%d['ms_per_step'], {k:round(v,4) for k,v in r.items() if k.startswith('k_drift')})"
done
