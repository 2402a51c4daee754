#!/bin/bash
mkdir -p gpurun_out
cp kmergutsjava_b200/libkmerguts_b200.so /tmp/default.so
run() { label=$1
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-legs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label', round(d['ms_per_step'],3), d['stage_ms'])"
}
run default
for v in U8 U2 OCC5 OCC7 OCC8 BLK256 BLK64; do
  cp kmergutsjava_b200/variants/lib_$v.so kmergutsjava_b200/libkmerguts_b200.so
  run $v
done
cp /tmp/default.so kmergutsjava_b200/libkmerguts_b200.so
