#!/bin/bash
mkdir -p gpurun_out
run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-legs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label', round(d['ms_per_step'],3), d['stage_ms'])"
}
run base KG_X=1
run carve72 KG_L2_CARVE_MB=72
run carve79 KG_L2_CARVE_MB=79
run f56_carve64 KG_FILTER_MAX_MB=56 KG_L2_CARVE_MB=64
run f56_carve79 KG_FILTER_MAX_MB=56 KG_L2_CARVE_MB=79
run f48_carve64 KG_FILTER_MAX_MB=48 KG_L2_CARVE_MB=64
run f70_carve79 KG_FILTER_MAX_MB=70 KG_L2_CARVE_MB=79
run hit09 KG_L2_HIT_RATIO=0.9
