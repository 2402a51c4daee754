#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_c0_ecoli.py -m gpu -x -q > gpurun_out/r14_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r14_pytest.log
timeout 600 python tests/configs/config2_dna.py --steps 10 --parity-genomes 50 > gpurun_out/r14_dna.json 2> gpurun_out/r14_dna.err; echo "dna rc=$?"; tail -1 gpurun_out/r14_dna.err
cut -c1-700 gpurun_out/r14_dna.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r14_dna_launches.csv python tests/configs/config2_dna.py --steps 2 --parity-genomes 0 > gpurun_out/r14_dna_ncu.log 2>&1; echo "rc=$?"
