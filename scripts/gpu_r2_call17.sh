#!/bin/bash
# round 2, call 17 (1 GPU): rsg_ram_run_host (upload | step | download pipelined over pitch-angle chunks) on hardware:
# parity with the three calls, and the e2e figure of the bench on both grids
set -u
O=gpurun_out/r2o
mkdir -p $O
timeout 600 python -m pytest tests/test_ram_parity_gpu.py -q -k "ram_run_host or nonperiodic" > $O/test_run_host.log 2>&1; tail -4 $O/test_run_host.log
B="--no-cpu-baseline --no-scb --no-extras --no-configs1"
for W in x4 default; do
timeout 300 python bench.py $B --steps 20 --workload $W > $O/bench_$W.json 2> $O/bench_$W.err
python - <<PY
import json
try:
    d = json.loads(open("$O/bench_$W.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("$W ms/step %.4f e2e ms %.3f (three calls %.3f) traffic %s" % (d["ms_per_step"], e["ms_per_step"], e.get("three_calls_ms_per_step", -1), d["roofline"]["traffic"]))
except Exception as ex:
    print("$W ERR", ex); print(open("$O/bench_$W.err").read()[-1500:])
PY
done
