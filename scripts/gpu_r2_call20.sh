#!/bin/bash
# round 2, call 20 (1 GPU): final state -- the whole -m gpu suite and the default bench line
set -u
O=gpurun_out/r2r
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu > $O/gpu_suite.log 2>&1; tail -4 $O/gpu_suite.log
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 200 $O/bench_n1.json; echo
