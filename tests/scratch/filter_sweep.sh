#!/bin/bash
# experiment: prefilter size vs k_probe time on configs[1] (filter larger than the 79 MiB persisting carve-out)
for cfg in "3 76" "3.5 88" "4 100" "4.5 112" "2.5 76"; do
  set -- $cfg
  KG_FILTER_BITS=$1 KG_FILTER_MAX_MB=$2 timeout 200 python bench.py --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/fs.out 2> gpurun_out/fs.err
  echo "bits $1 max $2 MB: $(grep '^{' gpurun_out/fs.out | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), d["stage_ms"])')"
done
