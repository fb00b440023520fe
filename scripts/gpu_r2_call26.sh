#!/bin/bash
# round 2, call 26 (1 GPU): cluster SOR vs one-CTA SOR on a grid 4 x the default SCB grid (configs[4])
set -u
mkdir -p gpurun_out/r2x
timeout 150 python -m pytest tests/test_scb_parity_gpu.py -q -k "cluster_matches and grid3" -s > gpurun_out/r2x/test_scb4x.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2x/test_scb4x.log
