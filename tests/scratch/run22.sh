#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | sed -n 2p
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r22_bench8.json 2> gpurun_out/r22_bench8.err; echo "bench8 rc=$?"; tail -4 gpurun_out/r22_bench8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r22_bench8.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e"):
    print(k, json.dumps(d.get(k))[:900])
for k in ("configs2","configs3","configs4"):
    print(k, json.dumps(d.get(k))[:1600])
PY
KG_SHARD_TRANSPORT=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tests/configs/config4_sharded.py --gpus 8 --steps 10 > gpurun_out/r22_c4_nccl.json 2> gpurun_out/r22_c4_nccl.err; echo "c4 nccl rc=$?"; tail -1 gpurun_out/r22_c4_nccl.json | cut -c1-1200
