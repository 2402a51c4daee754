#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "many or aa_parity" > gpurun_out/r28_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r28_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-legs 2>gpurun_out/r28_bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['value_one_call_at_a_time'], d['ms_per_step_one_call_at_a_time'], d['roofline'], d['stage_ms'])"
tail -3 gpurun_out/r28_bench.err
