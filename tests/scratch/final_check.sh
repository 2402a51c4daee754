#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
python -c "import json,sys; d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], [ (k, d[k].get('error') or d[k].get('parity') or d[k].get('parity_hits')) for k in ('configs2','configs3','configs4','table_load')])"
