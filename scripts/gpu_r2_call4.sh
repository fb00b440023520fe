#!/bin/bash
# round 2, fourth 1-GPU call: SCB with one-wave alpha clusters and the callback-free scb_run; hardware run of the new RAM<->SCB pieces
set -u
O=gpurun_out/r2b
mkdir -p $O
timeout 1500 python -m pytest tests/test_baseline_grids_gpu.py tests/test_scb_parity_gpu.py tests/test_zz_late_additions_gpu.py -q -s > $O/tests_scb.log 2>&1; grep -E "passed|failed|alpha:|SETRC|configs" $O/tests_scb.log | tail -12
python - > $O/scb_timing.log 2>&1 <<'PY'
import json, os, sys
sys.path.insert(0, os.getcwd())
import bench
for wide in ("0", "1"):
    if wide == "0": os.environ["RSG_SCB_NO_WIDE"] = "1"
    else: os.environ.pop("RSG_SCB_NO_WIDE", None)
    m = bench.scb_metrics(0)
    print("wide", wide, json.dumps({k: (m[k] if not isinstance(m[k], dict) else {q: m[k][q] for q in ("ms", "max_sweeps", "sweeps_per_s")}) for k in ("iterate_alpha", "iterate_psi", "bandjacob_ms", "metrica_ms")}))
    r = bench.scb_run_metrics(0)
    print("wide", wide, "scb_run", json.dumps({k: {q: r[k][q] for q in ("wall_ms", "outer_iterations", "launches", "ms_per_outer_iteration")} for k in ("device_front_end", "host_callback")}))
PY
cat $O/scb_timing.log | tail -6
export RSG_NO_GRAPH=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_scb_run.csv python -c "import bench, json; print(json.dumps(bench.scb_run_metrics(0)['device_front_end']))" > $O/launches_scb_run.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_computehI.csv python -c "import bench, json; print(json.dumps(bench.hi_metrics(0)))" > $O/launches_computehI.log 2>&1
ncu --set full --clock-control none -k 'regex:^(k_hi_nn9|k_hi_lines|k_hi_smooth|k_scb_sor_cluster_reg)' --launch-count 6 -o $O/full_misc -f python -c "import bench; bench.hi_metrics(0); bench.scb_metrics(0)" > $O/full_misc.log 2>&1
ncu -i $O/full_misc.ncu-rep --page raw --csv > $O/full_misc_raw.csv
rm -f $O/*.ncu-rep; du -sh $O
