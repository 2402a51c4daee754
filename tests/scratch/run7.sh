#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
export KG_PROBE2_KIND=own
for P in 4 5 6; do
KG_PROBE2_POLICY=$P timeout 300 $B > gpurun_out/r7_pol$P.json 2> gpurun_out/r7_pol$P.err; echo "pol $P rc=$?"
done
python - <<'PY'
import json
for n in ("4","5","6"):
    try:
        d=json.loads(open(f"gpurun_out/r7_pol{n}.json").read().strip().splitlines()[-1])
        print("policy",n, "ms/step", round(d["ms_per_step"],3), "stage", d["stage_ms"], "frac", d["roofline"]["frac"])
    except Exception as e:
        print(n, "failed", e)
PY
