#!/bin/bash
# round 2, twelfth call (8 GPUs, charged 8x): the chunked / overlapped bulk re-sharding (RSG_PEER_PUSH=2) against the kernels'
# own peer write-backs on the same box, per-stage times
set -u
O=gpurun_out/r2j
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
P=29580
run() {   # name, env...
  local name=$1; shift
  P=$((P+1))
  env "$@" timeout 300 $TR --nproc-per-node 8 --master-port $P bench.py --gpus 8 --steps 20 --warmup 3 --policy ${POL:-slabs} > $O/bench_n8_$name.json 2> $O/bench_n8_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/bench_n8_$name.json").read().strip().splitlines()[-1])
    ps = (d["roofline"].get("per_stage") or {}).get("max_over_ranks_ms", {})
    print("N=8 $name ms/step %.4f e2e ms %.3f check %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"].get("sharded_check", {}).get("every_rank_share_of_F2_bit_identical_to_one_gpu_step")))
    print("   ", {k: round(v, 4) for k, v in ps.items()})
except Exception as e:
    print("N=8 $name ERR", e)
PY
}
POL=slabs run slabs_push0 RSG_PEER_PUSH=0
POL=slabs run slabs_push2_c4 RSG_PEER_PUSH=2 RSG_PEER_CHUNKS=4
POL=slabs run slabs_push2_c2 RSG_PEER_PUSH=2 RSG_PEER_CHUNKS=2
POL=slabs run slabs_push2_c8 RSG_PEER_PUSH=2 RSG_PEER_CHUNKS=8
POL=species run species_push2_c4 RSG_PEER_PUSH=2 RSG_PEER_CHUNKS=4
