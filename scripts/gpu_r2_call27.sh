#!/bin/bash
set -u
mkdir -p gpurun_out/r2y
timeout 120 python scripts/time_configs4_size.py > gpurun_out/r2y/configs4_size_timings.json 2> gpurun_out/r2y/err.log; echo "rc=$?"; tail -3 gpurun_out/r2y/configs4_size_timings.json; tail -3 gpurun_out/r2y/err.log
