#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_shard.py -m gpu -x -q -k nccl > gpurun_out/r21_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r21_pytest.log
for tr in direct nccl; do
KG_SHARD_TRANSPORT=$tr timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/configs/config4_sharded.py --gpus 2 --steps 10 > gpurun_out/r21_c4_$tr.json 2> gpurun_out/r21_c4_$tr.err; echo "c4 $tr rc=$?"; tail -3 gpurun_out/r21_c4_$tr.err; tail -1 gpurun_out/r21_c4_$tr.json | cut -c1-1500
done
