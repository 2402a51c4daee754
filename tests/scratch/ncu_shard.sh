#!/bin/bash
# ncu evidence for the hash-sharded mode's kernels on one GPU (N = 1: no interconnect; 293 M signatures, 1 M proteins)
mkdir -p gpurun_out
S=tests/configs/config4_sharded.py
timeout 300 python $S --gpus 1 --steps 5 --warmup 3 --no-check > gpurun_out/r50_shard.json 2> gpurun_out/r50_shard.err; echo "plain rc=$?"; tail -c 600 gpurun_out/r50_shard.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r50_shard_launches.csv \
  python $S --gpus 1 --steps 1 --warmup 3 --no-check > gpurun_out/r50_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on \
  -k regex:'k_route|k_answer|k_mark_replies|k_word_popc|k_place_replies|k_tile_meta' --launch-skip 18 -c 6 \
  -o gpurun_out/r50_shard_kernels -f python $S --gpus 1 --steps 1 --warmup 3 --no-check > gpurun_out/r50_ncu2.log 2>&1; echo "full rc=$?"
tail -3 gpurun_out/r50_ncu2.log
ls -la gpurun_out | tail -8
