#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r33_bench8.json 2> gpurun_out/r33_bench8.err; echo "bench8 rc=$?"; tail -3 gpurun_out/r33_bench8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r33_bench8.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","two_in_flight","e2e","clocks"):
    print(k, json.dumps(d.get(k))[:700])
for k in ("configs2","configs3","configs4"):
    print(k, json.dumps(d.get(k))[:1300])
PY
