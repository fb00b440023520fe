#!/bin/bash
# round 2, ninth call (8 GPUs, charged 8x: kept short): where the N = 8 step spends its time (per-stage device times) and the
# bulk-push alternative for the second re-sharding (RSG_PEER_PUSH=1) against the column kernel's own peer write-back
set -u
O=gpurun_out/r2g
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
P=29560
for PUSH in 0 1; do
for POL in slabs species; do
P=$((P+1))
RSG_PEER_PUSH=$PUSH timeout 300 $TR --nproc-per-node 8 --master-port $P bench.py --gpus 8 --steps 20 --warmup 3 --policy $POL > $O/bench_n8_${POL}_push$PUSH.json 2> $O/bench_n8_${POL}_push$PUSH.err
python - <<PY
import json
try:
    d = json.loads(open("$O/bench_n8_${POL}_push$PUSH.json").read().strip().splitlines()[-1])
    ps = (d["roofline"].get("per_stage") or {}).get("max_over_ranks_ms", {})
    print("N=8 $POL push=$PUSH ms/step %.4f e2e ms %.3f check %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"].get("sharded_check", {}).get("every_rank_share_of_F2_bit_identical_to_one_gpu_step")))
    print("   ", {k: round(v, 4) for k, v in ps.items()})
except Exception as e:
    print("N=8 $POL push=$PUSH ERR", e)
PY
done
done
