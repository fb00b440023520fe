#!/bin/bash
# round 2, call 21 (2 GPUs): the multi-GPU tests of the suite with the final code (NCCL path, peer-memory path with one
# and two species pipelines, SCB sharding)
set -u
O=gpurun_out/r2s
mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu -k "multi_gpu" > $O/multi_gpu_tests.log 2>&1; tail -5 $O/multi_gpu_tests.log
