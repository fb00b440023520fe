#!/bin/bash
# Round-1 profiling pass (run on the GPU box through gpurun; outputs under gpurun_out/r1/).
#   1. launch list of the default bench command (kernel share of the step)
#   2. one `--set full` capture of each kernel of the fused step, default grid and 4x grid
#   3. the same for the one-kernel-per-operator FAST path (RSG_NO_FUSE=1), default grid
set -u
O=gpurun_out/r1
mkdir -p $O
export RSG_NO_GRAPH=1     # kernel-by-kernel launches so every launch is a separate ncu result
K='regex:^(k_plane_rp|k_col_fused|k_anisch_pa_fast|k_finalize)$'
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-scb > $O/launches_default.log 2>&1
for W in default x4; do
  ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 15 --launch-count 5 \
      -o $O/full_$W -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-scb --workload $W > $O/full_$W.log 2>&1
  ncu -i $O/full_$W.ncu-rep --page raw --csv > $O/full_${W}_raw.csv
done
# 4. SCB: launch list of one Euler-potential solve (bench.py's scb_metrics)
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_scb.csv \
    python scripts/scb_time.py > $O/launches_scb.log 2>&1
