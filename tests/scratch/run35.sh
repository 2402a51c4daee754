#!/bin/bash
mkdir -p gpurun_out
run() { label=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-legs --no-e2e2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('$label', 'packed', round(e['ms_per_step'],3), 'raw', round(e['raw_bytes_call']['ms_per_step'],3))"
}
run base KG_X=1
run mb40 KG_SLICE_MB=40
run mb64 KG_SLICE_MB=64
run mb80 KG_SLICE_MB=80
run ramp4 KG_SLICE_RAMP=4
run ramp16 KG_SLICE_RAMP=16
run mb64_ramp16 KG_SLICE_MB=64 KG_SLICE_RAMP=16
run mb80_ramp16 KG_SLICE_MB=80 KG_SLICE_RAMP=16
