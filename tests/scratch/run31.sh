#!/bin/bash
mkdir -p gpurun_out
# 4 virtual ranks on one GPU, 250k proteins each: the kernels see the bin / owner structure of a 4-GPU run
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r31_shard_launches.csv python tests/configs/config4_sharded.py --local-ranks 4 --families 1400000 --proteins 500000 --steps 2 --warmup 1 > gpurun_out/r31_shard.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r31_shard.log | cut -c1-600
