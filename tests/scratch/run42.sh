#!/bin/bash
mkdir -p gpurun_out
for tr in direct copy; do
KG_SHARD_TRANSPORT=$tr timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r42_$tr.csv python tests/configs/config4_sharded.py --local-ranks 4 --families 1400000 --proteins 500000 --steps 2 --warmup 1 > gpurun_out/r42_$tr.log 2>&1; echo "$tr rc=$?"
done
