#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/r3_plain.json 2> gpurun_out/r3_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:'k_filter|k_refilter|k_probe2' -s 9 -c 3 -o gpurun_out/r3_cascade $CMD > gpurun_out/r3_ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3_plain.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],3), "stage", d["stage_ms"])
PY
