#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_c0_ecoli.py -m gpu -x -q > gpurun_out/r36_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r36_pytest.log
run() { label=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-legs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('$label', 'value', round(d['ms_per_step'],3), 'two_in_flight', round(d['two_in_flight']['ms_per_step'],3), 'packed', round(e['ms_per_step'],3), 'raw', round(e['raw_bytes_call']['ms_per_step'],3), 'two_threads', round(e['two_threads']['ms_per_step'],3))"
}
run two_streams KG_X=1
run one_stream KG_ONE_COMPUTE_STREAM=1
run two_streams KG_X=1
run one_stream KG_ONE_COMPUTE_STREAM=1
