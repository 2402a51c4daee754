#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r10_dna_launches.csv python tests/configs/config2_dna.py --steps 2 --parity-genomes 0 > gpurun_out/r10_dna.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/r10_dna.log | cut -c1-600
