#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
nproc
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r9_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r9_pytest.log
timeout 200 python tests/scratch/first_call.py > gpurun_out/r9_first_call.log 2>&1; echo "first_call rc=$?"; tail -12 gpurun_out/r9_first_call.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r9_bench.json 2> gpurun_out/r9_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r9_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r9_bench.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","roofline","cpu_baseline","parity","stage_ms"):
    print(k, json.dumps(d.get(k))[:1200])
for k in ("configs2","configs3","configs4"):
    print(k, json.dumps(d.get(k))[:1500])
PY
