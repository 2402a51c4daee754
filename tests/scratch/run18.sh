#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r18_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r18_pytest.log
for m in seq; do
KG_FSM=$m timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-legs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m', d['ms_per_step'], d['stage_ms'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r18_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-legs > gpurun_out/r18_ncu.log 2>&1; echo "rc=$?"
