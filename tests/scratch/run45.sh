#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r45_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r45_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r45_bench.json 2> gpurun_out/r45_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r45_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r45_bench.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","two_in_flight","e2e","roofline","stage_ms","table_load","parity","parity_full_size","gpu_launches","clocks"):
    print(k, json.dumps(d.get(k))[:500])
for k in ("configs2","configs3","configs4"):
    print(k, json.dumps(d.get(k))[:400])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_translate|k_gather_seg|k_fsm_seg|k_otu_fold' -c 4 -o gpurun_out/r45_dna_kernels python tests/configs/config2_dna.py --steps 1 --parity-genomes 0 > gpurun_out/r45_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r45_dna_launches.csv python tests/configs/config2_dna.py --steps 2 --parity-genomes 0 > gpurun_out/r45_ncu2.log 2>&1; echo "ncu list rc=$?"
