#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r25_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r25_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r25_bench.json 2> gpurun_out/r25_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r25_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r25_bench.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","roofline","stage_ms","table_load"):
    print(k, json.dumps(d.get(k))[:700])
for k in ("configs2","configs3","configs4"):
    print(k, json.dumps(d.get(k))[:500])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r25_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-legs > gpurun_out/r25_ncu1.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_probe|k_fsm$' -c 2 -o gpurun_out/r25_k_probe python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-legs > gpurun_out/r25_ncu2.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r25_k_probe.ncu-rep --page raw --csv > gpurun_out/r25_k_probe_raw.csv 2>/dev/null
