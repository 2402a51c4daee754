#!/bin/bash
# experiment: slice plan of kg_run in 6-frame mode (configs[2] end to end)
for cfg in "20 8" "48 8" "48 2" "96 2" "96 1" "128 2" "256 1"; do
  set -- $cfg
  KG_SLICE_MB=$1 KG_SLICE_RAMP=$2 timeout 200 python tests/configs/config2_dna.py --steps 5 --parity-genomes 0 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('slice $1 MB ramp $2: e2e', round(d['e2e']['ms_per_step'],2), 'ms; device', round(d['ms_per_step'],2))"
done
