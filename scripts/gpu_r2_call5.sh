#!/bin/bash
# round 2, fifth call (2 GPUs): the whole -m gpu suite incl. the multi-GPU tests (RAM over NCCL and over peer memory, SCB sub-problem
# and zeta sharding), then scb_run with the warp-per-line maps, the zeta-sharded SOR on 2 GPUs timed
set -u
O=gpurun_out/r2c
mkdir -p $O
timeout 1800 python -m pytest tests -q -m gpu > $O/gpu_suite.log 2>&1; tail -6 $O/gpu_suite.log
python - > $O/scb_timing.log 2>&1 <<'PY'
import json, os, sys
sys.path.insert(0, os.getcwd())
import bench
for serial in ("1", "0"):
    if serial == "1": os.environ["RSG_SCB_MAP_SERIAL"] = "1"
    else: os.environ.pop("RSG_SCB_MAP_SERIAL", None)
    m = bench.scb_metrics(0)
    print("serial_maps", serial, json.dumps({k: m.get(k) for k in ("map_alpha_ms", "map_psi_ms", "map_theta_ms")}))
    r = bench.scb_run_metrics(0)
    print("serial_maps", serial, "scb_run", json.dumps({k: {q: r[k][q] for q in ("wall_ms", "outer_iterations", "launches", "ms_per_outer_iteration")} for k in ("device_front_end", "host_callback")}))
PY
cat $O/scb_timing.log | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu_scb_check.py > $O/multi_gpu_scb_check.log 2>&1; grep -v "^\*\|^$\|OMP_NUM" $O/multi_gpu_scb_check.log | tail -12
