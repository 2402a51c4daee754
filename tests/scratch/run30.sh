#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r30_bench2.json 2> gpurun_out/r30_bench2.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r30_bench2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r30_bench2.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","two_in_flight","e2e"):
    print(k, json.dumps(d.get(k))[:900])
for k in ("configs2","configs3","configs4"):
    print(k, json.dumps(d.get(k))[:900])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r30_ref2.json 2> gpurun_out/r30_ref2.err; echo "ref2 rc=$?"; tail -2 gpurun_out/r30_ref2.err; cut -c1-300 gpurun_out/r30_ref2.json
