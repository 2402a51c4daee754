#!/bin/bash
mkdir -p gpurun_out
for v in normal skiphi; do
  if [ $v = skiphi ]; then export KG_ROUTE_SKIP_HI=1; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/configs/config4_sharded.py --gpus 2 --steps 10 --no-check 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['ms_per_step'],3), d['rank0_phase_ms'])"
done
