#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --legs 4 > gpurun_out/r23_bench8.json 2> gpurun_out/r23_bench8.err; echo "bench8 rc=$?"; tail -3 gpurun_out/r23_bench8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r23_bench8.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e"):
    print(k, json.dumps(d.get(k))[:1400])
for k in ("configs4",):
    print(k, json.dumps(d.get(k))[:1600])
PY
