#!/bin/bash
# experiment: prefilter size cap vs k_answer on a 2.6e8-signature shard (one rank)
for mb in 76 64 52; do
  KG_FILTER_MAX_MB=$mb timeout 200 python tests/configs/config4_sharded.py --steps 5 --no-check 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('filter cap $mb MB:', round(d['ms_per_step'],2), d['rank0_phase_ms'])"
done
