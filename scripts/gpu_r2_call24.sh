#!/bin/bash
# round 2, call 24 (1 GPU): __graft_entry__.smoke() with the final code
set -u
mkdir -p gpurun_out/r2v
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/r2v/smoke.log
