#!/bin/bash
# round 2, 1-GPU call: in-process multi-rank tests of the sharded step, regression run of the GPU suite, first N=1 numbers
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_ram_shard_gpu.py -x -q 2>&1 | tail -25
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-scb --no-cpu-baseline --no-extras > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.err
python - <<'PY'
import json
for f in ("bench_n1",):
    try:
        l=[x for x in open(f"gpurun_out/{f}.json") if x.startswith("{")][-1]
        d=json.loads(l)
        print(f, "ms/step", d["ms_per_step"], "value", d["value"], "e2e ms", (d.get("e2e") or {}).get("ms_per_step"))
        if d.get("roofline"): print("   ", {k: d["roofline"].get(k) for k in ("kernel","frac","per_kernel_ms")})
        c=d.get("configs1") or {}
        print("configs1", c.get("ms_per_step"), c.get("value"), (c.get("e2e") or {}).get("ms_per_step"), (c.get("roofline") or {}).get("per_kernel_ms"))
    except Exception as e:
        print(f, "FAILED", e)
PY
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_ram_shard_gpu.py 2>&1 | tail -8
