#!/bin/bash
# round 2, call 18 (1 GPU): the unrolled ANISCH read loop -- the RAM parity tests, the step on both grids, and the ncu
# captures of the final kernels again (ram_kernels.cuh changed: traffic.json's sha)
set -u
O=gpurun_out/r2p
mkdir -p $O
timeout 900 python -m pytest tests/test_ram_parity_gpu.py tests/test_baseline_grids_gpu.py -q -x > $O/test_ram.log 2>&1; tail -4 $O/test_ram.log
B="--no-cpu-baseline --no-scb --no-extras --no-configs1"
for W in x4 default; do
timeout 300 python bench.py $B --steps 20 --workload $W > $O/bench_$W.json 2> $O/bench_$W.err
python - <<PY
import json
d = json.loads(open("$O/bench_$W.json").read().strip().splitlines()[-1])
print("$W ms/step %.4f e2e %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), {k: round(v, 4) for k, v in d["roofline"]["per_kernel_ms"].items()})
PY
done
export RSG_NO_GRAPH=1
K='regex:^(k_plane_rp|k_col_fused|k_wpadif_tables|k_anisch_pa_fast|k_finalize_wpi|k_finalize)$'
for W in x4 default; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_$W.csv python bench.py --steps 2 --warmup 3 $B --workload $W > $O/launches_$W.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 12 --launch-count 6 -o $O/full_$W -f python bench.py --steps 1 --warmup 3 $B --workload $W > $O/full_$W.log 2>&1
ncu -i $O/full_$W.ncu-rep --page raw --csv > $O/full_${W}_raw.csv
done
rm -f $O/*.ncu-rep; du -sh $O
