#!/bin/bash
# round 2, third 1-GPU call: the tests that failed / did not run in call 2, then the ncu evidence of the round:
# launch lists (shares) and --set full captures (DRAM bytes, stall reasons) of the fused RAM step on both grids, of scb_run and of computehI
set -u
O=gpurun_out/r2
mkdir -p $O
timeout 1200 python -m pytest tests/test_baseline_grids_gpu.py -q -s > $O/test_baseline_grids.log 2>&1; tail -5 $O/test_baseline_grids.log
timeout 600 python -m pytest tests/test_zz_late_additions_gpu.py tests/test_ram_parity_gpu.py -q -k "anisch or fused_wpadif or fast_mode_full" > $O/test_misc.log 2>&1; tail -4 $O/test_misc.log
export RSG_NO_GRAPH=1     # kernel-by-kernel launches so every launch is a separate ncu result
K='regex:^(k_plane_rp|k_col_fused|k_wpadif_tables|k_anisch_pa_fast|k_finalize_wpi|k_finalize)$'
B="--no-cpu-baseline --no-scb --no-extras --no-configs1"
for W in x4 default; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_$W.csv python bench.py --steps 2 --warmup 3 $B --workload $W > $O/launches_$W.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 12 --launch-count 6 -o $O/full_$W -f python bench.py --steps 1 --warmup 3 $B --workload $W > $O/full_$W.log 2>&1
ncu -i $O/full_$W.ncu-rep --page raw --csv > $O/full_${W}_raw.csv
ncu -i $O/full_$W.ncu-rep --page source --csv -k regex:k_col_fused > $O/full_${W}_source_col.csv 2>/dev/null; gzip -f $O/full_${W}_source_col.csv
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_scb_run.csv python -c "import bench, json; print(json.dumps(bench.scb_run_metrics(0)))" > $O/launches_scb_run.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_computehI.csv python -c "import bench, json; print(json.dumps(bench.hi_metrics(0)))" > $O/launches_computehI.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:^(k_hi_nn9|k_hi_lines|k_hi_smooth)' --launch-count 4 -o $O/full_computehI -f python -c "import bench; bench.hi_metrics(0)" > $O/full_computehI.log 2>&1
ncu -i $O/full_computehI.ncu-rep --page raw --csv > $O/full_computehI_raw.csv
# the reports themselves are too big to bring back (64 MiB limit on gpurun_out): the raw CSV pages carry every metric
rm -f $O/*.ncu-rep $O/*.ncu-rep.tmp; du -sh $O; ls -la $O | head -30
