#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tests/configs/config4_sharded.py --gpus 8 --steps 10 > gpurun_out/r24_c4_direct.json 2> gpurun_out/r24_c4_direct.err; echo "c4 direct rc=$?"; tail -1 gpurun_out/r24_c4_direct.json | cut -c1-1300
