#!/bin/bash
# round 2, call 15 (8 GPUs, charged 8x): the two-pipeline sharded step against the single pipeline on the same box
set -u
O=gpurun_out/r2m
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for PIPES in 2 1; do
RSG_PEER_PIPES=$PIPES RSG_BARRIER_TIMEOUT_MS=5000 timeout 200 $TR --nproc-per-node 8 --master-port 2962$PIPES bench.py --gpus 8 --steps 20 --warmup 3 > $O/bench_n8_slabs_pipes$PIPES.json 2> $O/bench_n8_slabs_pipes$PIPES.err
python - <<PY
import json
try:
    d = json.loads(open("$O/bench_n8_slabs_pipes$PIPES.json").read().strip().splitlines()[-1])
    print("N=8 slabs pipes=$PIPES ms/step %.4f e2e ms %.3f check %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"].get("sharded_check", {}).get("every_rank_share_of_F2_bit_identical_to_one_gpu_step")))
except Exception as e:
    print("pipes=$PIPES ERR", e)
PY
done
RSG_PEER_PIPES=2 RSG_BARRIER_TIMEOUT_MS=5000 timeout 200 $TR --nproc-per-node 4 --master-port 29631 bench.py --gpus 4 --steps 20 --warmup 3 > $O/bench_n4_slabs_pipes2.json 2> $O/bench_n4_slabs_pipes2.err
python - <<PY
import json
try:
    d = json.loads(open("$O/bench_n4_slabs_pipes2.json").read().strip().splitlines()[-1])
    print("N=4 slabs pipes=2 ms/step %.4f e2e ms %.3f check %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"].get("sharded_check", {}).get("every_rank_share_of_F2_bit_identical_to_one_gpu_step")))
except Exception as e:
    print("N=4 ERR", e)
PY
