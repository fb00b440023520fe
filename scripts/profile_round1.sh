#!/bin/bash
# Round-1 profiling pass (run on the GPU box through gpurun; outputs under gpurun_out/r1/).
#   1. launch list of the default bench command (kernel share of the step)
#   2. one `--set full` capture of each hot kernel, default grid and 4x grid
set -u
O=gpurun_out/r1
mkdir -p $O
export RSG_NO_GRAPH=1     # kernel-by-kernel launches so every launch is a separate ncu result
K='regex:^(k_driftr|k_driftp|k_drifte|k_driftmu|k_loss_mid|k_anisch_pa_fast|k_anisch_en)$'
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_default.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches_default.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 22 --launch-count 11 \
    -o $O/full_default -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/full_default.log 2>&1
ncu -i $O/full_default.ncu-rep --page raw --csv > $O/full_default_raw.csv
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 22 --launch-count 11 \
    -o $O/full_x4 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload x4 > $O/full_x4.log 2>&1
ncu -i $O/full_x4.ncu-rep --page raw --csv > $O/full_x4_raw.csv
