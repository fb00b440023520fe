#!/bin/bash
set -u
mkdir -p gpurun_out/r2z
timeout 100 python scripts/time_coupled_cycle.py > gpurun_out/r2z/coupled_cycle.json 2> gpurun_out/r2z/err.log; echo "rc=$?"; tail -2 gpurun_out/r2z/coupled_cycle.json; tail -3 gpurun_out/r2z/err.log
