#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/scratch/e2e_debug.py 2> gpurun_out/r32_e2e_debug.log; tail -14 gpurun_out/r32_e2e_debug.log
