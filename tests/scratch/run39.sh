#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_c0_ecoli.py tests/test_gpu_shard.py -m gpu -x -q > gpurun_out/r39_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r39_pytest.log
timeout 600 python tests/configs/config2_dna.py --steps 10 --parity-genomes 10 > gpurun_out/r39_dna.json 2> gpurun_out/r39_dna.err; echo "dna rc=$?"; tail -1 gpurun_out/r39_dna.err; cut -c1-800 gpurun_out/r39_dna.json
