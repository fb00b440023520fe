#!/bin/bash
# round 2, seventh call (8 GPUs, charged 8x: kept short): the library's sharded RAM step at N = 8 and N = 4, both shard
# policies, checked in-run against the one-GPU step; the peer-memory check script at 8 ranks on the default grid
set -u
O=gpurun_out/r2e
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
P=29540
for N in 8 4; do
for POL in slabs species; do
P=$((P+1))
timeout 300 $TR --nproc-per-node $N --master-port $P bench.py --gpus $N --steps 20 --warmup 3 --policy $POL > $O/bench_n${N}_$POL.json 2> $O/bench_n${N}_$POL.err
python - <<PY
import json
try:
    d = json.loads(open("$O/bench_n${N}_$POL.json").read().strip().splitlines()[-1])
    print("N=$N $POL ms/step %.4f value %.4g e2e ms %.3f check %s" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["config"].get("sharded_check", {}).get("every_rank_share_of_F2_bit_identical_to_one_gpu_step")))
except Exception as e:
    print("N=$N $POL ERR", e)
PY
done
done
P=$((P+1))
timeout 300 $TR --nproc-per-node 8 --master-port $P tests/multi_gpu_peer_check.py > $O/peer_check_n8.log 2>&1; grep -v "^\*\|^$\|OMP_NUM" $O/peer_check_n8.log | tail -6
