#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synth.py -m gpu -x -q > gpurun_out/r5_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r5_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
for U in 2 4 8; do
KG_PROBE2_U=$U timeout 300 $B > gpurun_out/r5_u$U.json 2> gpurun_out/r5_u$U.err; echo "U $U rc=$?"
done
KG_PROBE2_U=8 KG_CASCADE_PARTS=4 timeout 300 $B > gpurun_out/r5_u8p4.json 2> gpurun_out/r5_u8p4.err
python - <<'PY'
import json
for n in ("u2","u4","u8","u8p4"):
    try:
        d=json.loads(open(f"gpurun_out/r5_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],3), "stage", d["stage_ms"], "frac", d["roofline"]["frac"])
    except Exception as e:
        print(n, "failed", e)
PY
