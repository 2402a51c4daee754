#!/bin/bash
mkdir -p gpurun_out
KG_SHARD_TRANSPORT=direct timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_route' -c 1 -o gpurun_out/r44_route python tests/configs/config4_sharded.py --local-ranks 4 --families 1400000 --proteins 500000 --steps 1 --warmup 1 > gpurun_out/r44.log 2>&1; echo "rc=$?"
