#!/bin/bash
# round 2, eighth call (1 GPU): the sharded step with the bulk column push (in-process peers), the resident computehI on
# hardware (test + timing + ncu of its kernels), the whole late-additions file
set -u
O=gpurun_out/r2f
mkdir -p $O
timeout 900 python -m pytest tests/test_ram_shard_gpu.py -q > $O/test_shard.log 2>&1; tail -4 $O/test_shard.log
timeout 900 python -m pytest tests/test_zz_late_additions_gpu.py -q -s > $O/test_late.log 2>&1; tail -4 $O/test_late.log
python - > $O/hi_metrics.json 2> $O/hi_metrics.err <<'PY'
import json, os, sys
sys.path.insert(0, os.getcwd())
import bench
print(json.dumps(bench.hi_metrics(0)))
PY
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2f/hi_metrics.json"))
print(json.dumps({k: d[k] for k in ("convert_lines_default", "computehI_resident_default") if k in d}, indent=1))
PY
ncu --set full --clock-control none --import-source on -k 'regex:^(k_hi_nn9|k_hi_lines|k_hi_smooth|k_hi_tail_lines|k_hi_tail_cols|k_hi_rairden)' --launch-skip 30 --launch-count 12 -o $O/full_computehI -f python -c "import bench; bench.hi_metrics(0)" > $O/full_computehI.log 2>&1
ncu -i $O/full_computehI.ncu-rep --page raw --csv > $O/full_computehI_raw.csv
rm -f $O/*.ncu-rep; du -sh $O
