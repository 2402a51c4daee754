#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r8_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r8_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r8_bench.json 2> gpurun_out/r8_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r8_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r8_ref.json 2> gpurun_out/r8_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r8_bench.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","roofline","cpu_baseline","parity","stage_ms"):
    print(k, json.dumps(d.get(k))[:900])
for k in ("configs2","configs3","configs4"):
    print(k, json.dumps(d.get(k))[:1500])
PY
