#!/bin/bash
# `--set full` capture of the fused kernels (one step after warm-up), default grid and 4x grid.
set -u
TAG=${1:-fused}
O=gpurun_out/$TAG
mkdir -p $O
export RSG_NO_GRAPH=1
K='regex:^(k_plane_rp|k_col_fused)$'
for W in default x4; do
  ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 9 --launch-count 3 \
      -o $O/full_$W -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload $W > $O/full_$W.log 2>&1
  ncu -i $O/full_$W.ncu-rep --page raw --csv > $O/full_${W}_raw.csv
done
