#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r47_bench4.json 2> gpurun_out/r47_bench4.err; echo "bench4 rc=$?"; tail -2 gpurun_out/r47_bench4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r47_bench4.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["copy_only_ms_per_step"])
for k in ("configs2","configs3","configs4"):
    c=d[k]; print(k, c.get("ms_per_step"), c.get("lookups_per_s"), c.get("run_seconds"), (c.get("rank0_phase_ms") or ""))
PY
