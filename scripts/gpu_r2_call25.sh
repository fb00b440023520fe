#!/bin/bash
# round 2, call 25 (1 GPU): the property test at the size of configs[4]'s RAM grid (156 M cells)
set -u
mkdir -p gpurun_out/r2w
timeout 200 python -m pytest tests/test_baseline_grids_gpu.py -q -k "configs4" -s > gpurun_out/r2w/test_configs4.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2w/test_configs4.log
