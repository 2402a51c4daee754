#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_c0_ecoli.py -m gpu -x -q > gpurun_out/r29_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r29_pytest.log
timeout 600 python tests/configs/config2_dna.py --steps 10 --parity-genomes 0 > gpurun_out/r29_dna.json 2> gpurun_out/r29_dna.err; echo "dna rc=$?"; tail -2 gpurun_out/r29_dna.err; cut -c1-1500 gpurun_out/r29_dna.json
