#!/bin/bash
mkdir -p gpurun_out
KG_DEBUG_SEG=1 timeout 600 python tests/configs/config2_dna.py --steps 1 --parity-genomes 0 2>&1 | grep "kg seg" | head -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fsm_seg|k_otu_fold|k_gather_seg' -c 3 -o gpurun_out/r13_dna_group python tests/configs/config2_dna.py --steps 1 --parity-genomes 0 > gpurun_out/r13_ncu.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/r13_dna_group.ncu-rep --page raw --csv > gpurun_out/r13_dna_group_raw.csv 2>/dev/null; ls -la gpurun_out/ | head
