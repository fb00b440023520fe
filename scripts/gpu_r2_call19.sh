#!/bin/bash
# round 2, call 19 (1 GPU): pitch angles per thread of the ANISCH sums (RSG_ANISCH_LCH: launch shape only, no source change)
set -u
O=gpurun_out/r2q
mkdir -p $O
B="--no-cpu-baseline --no-scb --no-extras --no-configs1 --steps 10"
for W in x4 default; do
for LCH in 4 6 9 12 18 24; do
RSG_ANISCH_LCH=$LCH timeout 200 python bench.py $B --workload $W > $O/b_${W}_$LCH.json 2> $O/b_${W}_$LCH.err
python - <<PY
import json
try:
    d = json.loads(open("$O/b_${W}_$LCH.json").read().strip().splitlines()[-1])
    print("$W LCH=$LCH ms/step %.4f k_anisch %.4f" % (d["ms_per_step"], d["roofline"]["per_kernel_ms"]["k_anisch"]))
except Exception as e:
    print("$W $LCH ERR", e)
PY
done
done
