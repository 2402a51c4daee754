#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shard.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/r20_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r20_pytest.log
