#!/bin/bash
# round 2, call 14 (2 GPUs): the two-pipeline sharded step -- in-process peers with enough hardware queues, then one
# process per GPU over CUDA IPC (check script + bench, pipes 1 vs 2)
set -u
O=gpurun_out/r2l
mkdir -p $O
timeout 300 python -m pytest tests/test_ram_shard_gpu.py -q -x -k "pipelines" > $O/test_pipes.log 2>&1; tail -4 $O/test_pipes.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
RSG_PEER_PIPES=2 RSG_BARRIER_TIMEOUT_MS=5000 timeout 300 $TR --nproc-per-node 2 --master-port 29601 tests/multi_gpu_peer_check.py > $O/peer_check_n2_pipes2.log 2>&1; grep -v "^\*\|^$\|OMP_NUM" $O/peer_check_n2_pipes2.log | tail -5
for PIPES in 1 2; do
RSG_PEER_PIPES=$PIPES RSG_BARRIER_TIMEOUT_MS=5000 timeout 300 $TR --nproc-per-node 2 --master-port 2961$PIPES bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2_pipes$PIPES.json 2> $O/bench_n2_pipes$PIPES.err
python - <<PY
import json
try:
    d = json.loads(open("$O/bench_n2_pipes$PIPES.json").read().strip().splitlines()[-1])
    print("N=2 slabs pipes=$PIPES ms/step %.4f e2e ms %.3f check %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"].get("sharded_check", {}).get("every_rank_share_of_F2_bit_identical_to_one_gpu_step")))
except Exception as e:
    print("pipes=$PIPES ERR", e)
PY
done
