#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shard.py -m gpu -x -q > gpurun_out/r43_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r43_pytest.log
KG_SHARD_TRANSPORT=direct timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r43_direct.csv python tests/configs/config4_sharded.py --local-ranks 4 --families 1400000 --proteins 500000 --steps 2 --warmup 1 > gpurun_out/r43_direct.log 2>&1; echo "rc=$?"
