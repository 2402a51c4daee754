#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
export KG_PROBE2_U=4
$CMD > gpurun_out/r6_plain.json 2> gpurun_out/r6_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:'k_probe2' -s 3 -c 1 -o gpurun_out/r6_probe2 $CMD > gpurun_out/r6_ncu.log 2>&1
echo "ncu rc=$?"
