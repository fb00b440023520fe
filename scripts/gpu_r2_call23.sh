#!/bin/bash
# round 2, call 23 (1 GPU): racecheck over the SCB cluster SOR and the computehI kernels (small grids)
set -u
O=gpurun_out/r2u
mkdir -p $O
timeout 150 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_small.py scb > $O/racecheck_scb_hi.log 2>&1; echo "racecheck rc=$?"; tail -5 $O/racecheck_scb_hi.log
