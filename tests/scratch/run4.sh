#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synth.py -m gpu -x -q > gpurun_out/r4_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r4_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
for P in 1 2 4 8 16; do
KG_CASCADE_PARTS=$P timeout 300 $B > gpurun_out/r4_parts$P.json 2> gpurun_out/r4_parts$P.err; echo "parts $P rc=$?"
done
KG_CASCADE_PERSIST=1 KG_CASCADE_PARTS=4 timeout 300 $B > gpurun_out/r4_parts4p.json 2> gpurun_out/r4_parts4p.err
KG_CASCADE_PERSIST=1 KG_CASCADE_PARTS=8 timeout 300 $B > gpurun_out/r4_parts8p.json 2> gpurun_out/r4_parts8p.err
python - <<'PY'
import json
for n in ("1","2","4","8","16","4p","8p"):
    try:
        d=json.loads(open(f"gpurun_out/r4_parts{n}.json").read().strip().splitlines()[-1])
        print("parts",n, "ms/step", round(d["ms_per_step"],3), "stage", d["stage_ms"], "frac", d["roofline"]["frac"])
    except Exception as e:
        print(n, "failed", e)
PY
