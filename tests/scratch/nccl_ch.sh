#!/bin/bash
# experiment: NCCL point-to-point channel count vs exchange time of the sharded table (2 GPUs)
p=29520
for ch in 32 64; do
  p=$((p+1))
  NCCL_MIN_P2P_NCHANNELS=$ch NCCL_MAX_P2P_NCHANNELS=$ch timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $p tests/configs/config4_sharded.py --gpus 2 --steps 5 --no-check > gpurun_out/ch$ch.out 2> gpurun_out/ch$ch.err
  echo "ch $ch: $(grep '^{' gpurun_out/ch$ch.out | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],2), d["rank0_phase_ms"], d["interconnect_GBps_per_gpu_during_exchange"])')"
done
