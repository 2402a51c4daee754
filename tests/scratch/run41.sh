#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shard.py -m gpu -x -q -k nccl > gpurun_out/r41_pytest.log 2>&1; echo "pytest rc=$?"; tail -20 gpurun_out/r41_pytest.log
for tr in direct nccl; do
KG_SHARD_TRANSPORT=$tr timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/configs/config4_sharded.py --gpus 2 --steps 10 2>gpurun_out/r41_c4_$tr.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tr', round(d['ms_per_step'],3), d['rank0_phase_ms'], d['parity_hits'][:60], d['parity_calls'][:50])"
done
