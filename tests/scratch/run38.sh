#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shard.py -m gpu -x -q -k nccl > gpurun_out/r38_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r38_pytest.log
