#!/bin/bash
mkdir -p gpurun_out
free -g | head -2; df -h /dev/shm /tmp | tail -2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r19_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r19_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-legs > gpurun_out/r19_bench.json 2> gpurun_out/r19_bench.err; echo "bench rc=$?"; tail -6 gpurun_out/r19_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r19_bench.json").read().strip().splitlines()[-1])
print(json.dumps(d.get("table_load")))
PY
