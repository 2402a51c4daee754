#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "probe_variants" > gpurun_out/r34_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r34_pytest.log
run() { label=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-legs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label', round(d['ms_per_step'],3), d['stage_ms'], d['detail']['hits_per_step'])"
}
run base KG_X=1
run halves KG_FILTER_HALVES=1
run halves_bits4 KG_FILTER_HALVES=1 KG_FILTER_BITS=2
