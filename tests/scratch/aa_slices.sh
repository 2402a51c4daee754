#!/bin/bash
# experiment: slice plan of kg_run in protein mode (configs[1] end to end)
for cfg in "0 8" "52 4" "52 16" "40 8" "64 8" "32 4" "80 16"; do
  set -- $cfg
  if [ "$1" = "0" ]; then unset KG_SLICE_MB; else export KG_SLICE_MB=$1; fi
  KG_SLICE_RAMP=$2 timeout 200 python bench.py --steps 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('slice $1 MB ramp $2: e2e', round(d['e2e']['ms_per_step'],2), 'ms; device', round(d['ms_per_step'],2))"
done
