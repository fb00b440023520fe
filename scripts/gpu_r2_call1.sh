#!/bin/bash
# round 2, first GPU call (2 GPUs): the library's own multi-GPU step -- in-process ranks on one device, then one
# process per GPU over CUDA IPC -- and the first strong-scaling numbers of the configs[2] workload
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m 2>/dev/null | head -8
timeout 900 python -m pytest tests/test_ram_shard_gpu.py -x -q 2>&1 | tail -25
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 tests/multi_gpu_peer_check.py x4 2>&1 | tail -12
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-scb --no-cpu-baseline --no-extras > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.err
for pol in species slabs; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 --policy $pol > gpurun_out/bench_n2_$pol.json 2> gpurun_out/bench_n2_$pol.err; tail -c 600 gpurun_out/bench_n2_$pol.err
done
python - <<'PY'
import json
for f in ("bench_n1","bench_n2_species","bench_n2_slabs"):
    try:
        l=[x for x in open(f"gpurun_out/{f}.json") if x.startswith("{")][-1]
        d=json.loads(l)
        print(f, "ms/step", d["ms_per_step"], "value", d["value"], "e2e ms", (d.get("e2e") or {}).get("ms_per_step"), d["config"].get("sharded_check"))
        if d.get("roofline"): print("   ", {k: d["roofline"].get(k) for k in ("kernel","frac","per_kernel_ms","nvlink")})
    except Exception as e:
        print(f, "FAILED", e)
PY
