#!/bin/bash
# round 2, sixteenth call (1 GPU): the whole -m gpu suite with the final kernels, the default
# bench line, and the ncu captures of the final kernels (launch lists + --set full on both grids -> profiles/traffic.json)
set -u
O=gpurun_out/r2n
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu > $O/gpu_suite.log 2>&1; tail -5 $O/gpu_suite.log
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/bench_n1.json; echo
B="--no-cpu-baseline --no-scb --no-extras --no-configs1"
export RSG_NO_GRAPH=1
K='regex:^(k_plane_rp|k_col_fused|k_wpadif_tables|k_anisch_pa_fast|k_finalize_wpi|k_finalize)$'
for W in x4 default; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_$W.csv python bench.py --steps 2 --warmup 3 $B --workload $W > $O/launches_$W.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 12 --launch-count 6 -o $O/full_$W -f python bench.py --steps 1 --warmup 3 $B --workload $W > $O/full_$W.log 2>&1
ncu -i $O/full_$W.ncu-rep --page raw --csv > $O/full_${W}_raw.csv
done
rm -f $O/*.ncu-rep; du -sh $O
