#!/bin/bash
# experiment: chunked (overlapped) exchange of the sharded table on 2 GPUs

p=29540
for ch in 2 4; do
  p=$((p+1))
  KG_SHARD_CHUNKS=$ch timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $p tests/configs/config4_sharded.py --gpus 2 --steps 5 > gpurun_out/chk$ch.out 2> gpurun_out/chk$ch.err
  echo "chunks $ch: rc=$? $(grep '^{' gpurun_out/chk$ch.out | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],2), d["rank0_phase_ms"], d.get("parity_hits","")[:40])')"
done
