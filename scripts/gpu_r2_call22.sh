#!/bin/bash
# round 2, call 22 (1 GPU): compute-sanitizer over the hot kernels on small grids (memcheck on everything, racecheck on
# the shared-memory kernels of the RAM step)
set -u
O=gpurun_out/r2t
mkdir -p $O
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > $O/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/memcheck.log
timeout 120 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_small.py ram > $O/racecheck_ram.log 2>&1; echo "racecheck rc=$?"; tail -4 $O/racecheck_ram.log
