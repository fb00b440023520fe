#!/bin/bash
# round 2, sixth call (1 GPU): Coulomb stages of the fused column kernel on hardware (parity + sharded), the default bench
# line, flag-set timings, then the ncu evidence again for the kernels as they are now (launch lists + --set full on both grids,
# the three computehI kernels)
set -u
O=gpurun_out/r2d
mkdir -p $O
timeout 900 python -m pytest tests/test_ram_parity_gpu.py tests/test_ram_shard_gpu.py -q -s -k "coulomb" > $O/test_coulomb.log 2>&1; tail -6 $O/test_coulomb.log
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 600 $O/bench_n1.json
B="--no-cpu-baseline --no-scb --no-extras --no-configs1"
for F in 0 2 7; do
timeout 300 python bench.py $B --workload default --flags $F --steps 20 > $O/bench_default_f$F.json 2> $O/bench_default_f$F.err
timeout 300 python bench.py $B --workload x4 --flags $F --steps 10 > $O/bench_x4_f$F.json 2> $O/bench_x4_f$F.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2d/bench_*_f*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.4f launches/step %.1f" % (d["ms_per_step"], d["gpu_launches"] / d["steps"]), {k: round(v, 4) for k, v in (d.get("roofline") or {}).get("per_kernel_ms", {}).items()})
    except Exception as e:
        print(f, "ERR", e)
PY
export RSG_NO_GRAPH=1     # kernel-by-kernel launches so every launch is a separate ncu result
K='regex:^(k_plane_rp|k_col_fused|k_wpadif_tables|k_anisch_pa_fast|k_finalize_wpi|k_finalize)$'
for W in x4 default; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_$W.csv python bench.py --steps 2 --warmup 3 $B --workload $W > $O/launches_$W.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 12 --launch-count 6 -o $O/full_$W -f python bench.py --steps 1 --warmup 3 $B --workload $W > $O/full_$W.log 2>&1
ncu -i $O/full_$W.ncu-rep --page raw --csv > $O/full_${W}_raw.csv
done
ncu --set full --clock-control none --import-source on -k 'regex:^(k_hi_nn9|k_hi_lines|k_hi_smooth|k_hi_tail_lines)' --launch-count 5 -o $O/full_computehI -f python -c "import bench; bench.hi_metrics(0)" > $O/full_computehI.log 2>&1
ncu -i $O/full_computehI.ncu-rep --page raw --csv > $O/full_computehI_raw.csv
rm -f $O/*.ncu-rep $O/*.ncu-rep.tmp; du -sh $O
