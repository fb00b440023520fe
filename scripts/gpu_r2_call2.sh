#!/bin/bash
# round 2, second 1-GPU call: new tests on hardware, effect of the TMA staging / ANISCH fold / WPADIF pipelining, launch-shape sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_baseline_grids_gpu.py tests/test_ram_shard_gpu.py -x -q -s 2>&1 | grep -v "^$" | tail -15
timeout 900 python -m pytest tests/test_ram_parity_gpu.py tests/test_zz_late_additions_gpu.py -x -q -k "fast or fused or anisch or graph" 2>&1 | tail -5
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-scb --no-cpu-baseline --no-extras --no-configs1 ${WL} > gpurun_out/b_$name.json 2>gpurun_out/b_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([x for x in open(f"gpurun_out/b_{n}.json") if x.startswith("{")][-1])
    pk=d["roofline"]["per_kernel_ms"]
    print(f"{n:28s} ms/step {d['ms_per_step']:.4f}  " + " ".join(f"{k}={v:.4f}" for k,v in pk.items() if k.startswith("k_")), "exact", (d["config"].get("modes") or {}).get("exact_ms_per_step"))
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/b_{n}.err").read()[-300:])
PY
}
WL=""
run x4_base A=1
run x4_notma RSG_PLANE_TMA=0
run x4_nofold RSG_NO_ANISCH_FOLD=1
for T in 288 448 576 736 864 1024; do run x4_colT$T RSG_COL_T=$T; done
for KC in 2 4 6; do run x4_kc$KC RSG_KC_PLANE=$KC; done
for PT in 256 384; do run x4_pT$PT RSG_PLANE_T=$PT; done
WL="--workload default"
run d_base A=1
run d_notma RSG_PLANE_TMA=0
for T in 160 288 320 448 576; do run d_colT$T RSG_COL_T=$T; done
for KC in 4 6 8; do run d_kc$KC RSG_KC_PLANE=$KC; done
for PT in 192 256 384 512; do run d_pT$PT RSG_PLANE_T=$PT; done
