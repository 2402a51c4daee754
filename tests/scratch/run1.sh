#!/bin/bash
# first GPU pass of round 2: tests, then the bench with the cascade and with the fused probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -5 gpurun_out/r2_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cascade.json 2> gpurun_out/r2_bench_cascade.err; echo "cascade rc=$?"
KG_PROBE=fused timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_fused.json 2> gpurun_out/r2_bench_fused.err; echo "fused rc=$?"
KG_NO_L2_PERSIST=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_cascade_nopersist.json 2> gpurun_out/r2_bench_cascade_nopersist.err; echo "nopersist rc=$?"
python - <<'PY'
import json
for n in ("cascade","fused","cascade_nopersist"):
    try:
        d=json.loads(open(f"gpurun_out/r2_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],3), "stage", d["stage_ms"], "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],3), "frac", d["roofline"]["frac"])
    except Exception as e:
        print(n, "failed", e)
PY
