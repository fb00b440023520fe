#!/bin/bash
# First GPU call of round 2 (run on the GPU box through gpurun; outputs under gpurun_out/r2/): everything that was
# written after round 1's GPU minutes ran out gets its hardware run, timing and ncu capture in one go.
#   0. the -m gpu tests that have never run on hardware
#   1. bench line with the extras (WPI/EMIC step on both grids = configs[2], zeta protocol, scb_run = configs[3])
#   2. launch list + `--set full` capture of the WPI/EMIC step on the 4x grid (configs[2])
#   3. launch list of one whole scb_run (rsg_scb_run, default SCB grid)
set -u
O=gpurun_out/r2
mkdir -p $O
timeout 900 python -m pytest tests/test_zz_late_additions_gpu.py -m gpu -q > $O/late_additions_tests.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err
export RSG_NO_GRAPH=1     # kernel-by-kernel launches so every launch is a separate ncu result
K='regex:^(k_plane_rp|k_col_fused|k_wpadif_tables|k_anisch_pa_fast|k_finalize_wpi|k_finalize)$'
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_x4_wpi.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-scb --no-extras --workload x4 --flags 5 > $O/launches_x4_wpi.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 18 --launch-count 6 \
    -o $O/full_x4_wpi -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-scb --no-extras --workload x4 --flags 5 > $O/full_x4_wpi.log 2>&1
ncu -i $O/full_x4_wpi.ncu-rep --page raw --csv > $O/full_x4_wpi_raw.csv
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_scb_run.csv \
    python -c "import bench, json; print(json.dumps(bench.scb_run_metrics(0)))" > $O/launches_scb_run.log 2>&1
#   4. computehI (rsg_hI_convert_lines / rsg_hI_integrals / rsg_hI_tail): launch list and a --set full capture of the
#      nearest-neighbour kernel and of the integral kernel (both have never run on hardware before step 0)
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_computehI.csv \
    python -c "import bench, json; print(json.dumps(bench.hi_metrics(0)))" > $O/launches_computehI.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:^(k_hi_nn9|k_hi_lines|k_hi_smooth)' --launch-count 4 \
    -o $O/full_computehI -f python -c "import bench; bench.hi_metrics(0)" > $O/full_computehI.log 2>&1
ncu -i $O/full_computehI.ncu-rep --page raw --csv > $O/full_computehI_raw.csv
