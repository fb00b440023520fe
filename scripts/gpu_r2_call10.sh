#!/bin/bash
# round 2, tenth call (1 GPU): the chunked / overlapped bulk re-sharding with in-process peers, the N = 1 bench line after
# the column kernel's 576-thread register bucket, and the ncu captures again (ram_fused.cuh changed: traffic.json's sha)
set -u
O=gpurun_out/r2h
mkdir -p $O
timeout 900 python -m pytest tests/test_ram_shard_gpu.py tests/test_ram_parity_gpu.py -q -x -k "shard or fused or fast_mode_full or graph" > $O/test_shard.log 2>&1; tail -4 $O/test_shard.log
B="--no-cpu-baseline --no-scb --no-extras --no-configs1"
timeout 300 python bench.py $B --steps 20 > $O/bench_x4.json 2> $O/bench_x4.err
timeout 300 python bench.py $B --steps 20 --workload default > $O/bench_default.json 2> $O/bench_default.err
python - <<'PY'
import json
for f in ("gpurun_out/r2h/bench_x4.json", "gpurun_out/r2h/bench_default.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "ms/step %.4f" % d["ms_per_step"], {k: round(v, 4) for k, v in d["roofline"]["per_kernel_ms"].items()}, "traffic", d["roofline"]["traffic"])
PY
export RSG_NO_GRAPH=1
K='regex:^(k_plane_rp|k_col_fused|k_wpadif_tables|k_anisch_pa_fast|k_finalize_wpi|k_finalize)$'
for W in x4 default; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_$W.csv python bench.py --steps 2 --warmup 3 $B --workload $W > $O/launches_$W.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 12 --launch-count 6 -o $O/full_$W -f python bench.py --steps 1 --warmup 3 $B --workload $W > $O/full_$W.log 2>&1
ncu -i $O/full_$W.ncu-rep --page raw --csv > $O/full_${W}_raw.csv
done
rm -f $O/*.ncu-rep; du -sh $O
