#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fsm_seg' -c 1 -o gpurun_out/r16_fsm_seg python tests/configs/config2_dna.py --steps 1 --parity-genomes 0 > gpurun_out/r16_ncu.log 2>&1; echo "rc=$?"
