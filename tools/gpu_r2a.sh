#!/bin/bash
# round 2, call A: TMA Viterbi kernel -- parity tests, then timing against the per-row bulk-copy kernel; MN-major UMMA probe
mkdir -p gpurun_out
timeout 60 tools/ubench/mn_major_test 2>&1 | tail -6
( time timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -x -q -k "viterbi or emission" ) 2>&1 | tail -15 > gpurun_out/tests_a.log; tail -8 gpurun_out/tests_a.log
SAPR_TMA=0 timeout 120 python tools/vit_bench.py 100000 10 2>&1 | tail -1
timeout 120 python tools/vit_bench.py 100000 10 2>&1 | tail -1
timeout 120 python tools/vit_bench.py 94720 10 2>&1 | tail -1
