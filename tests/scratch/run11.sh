#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r12_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r12_pytest.log
timeout 600 python tests/configs/config2_dna.py --steps 10 --parity-genomes 50 > gpurun_out/r12_dna.json 2> gpurun_out/r12_dna.err; echo "dna rc=$?"; tail -3 gpurun_out/r12_dna.err
cut -c1-1200 gpurun_out/r12_dna.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r12_dna_launches.csv python tests/configs/config2_dna.py --steps 2 --parity-genomes 0 > gpurun_out/r12_dna_ncu.log 2>&1; echo "rc=$?"
