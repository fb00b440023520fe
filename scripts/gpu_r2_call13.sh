#!/bin/bash
# round 2, call 13 (1 GPU): the two-pipeline sharded step with in-process peers (correctness before the 8-GPU timing)
set -u
O=gpurun_out/r2k
mkdir -p $O
timeout 900 python -m pytest tests/test_ram_shard_gpu.py -q > $O/test_shard.log 2>&1; tail -6 $O/test_shard.log
