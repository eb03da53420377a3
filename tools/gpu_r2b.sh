#!/bin/bash
# round 2, call B: TMA Viterbi + fused grouped E-step -- parity tests, then timing against the round-1 kernels
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -x -q -k "viterbi or emission or grouped" ) 2>&1 | tail -25 > gpurun_out/tests_b.log; tail -8 gpurun_out/tests_b.log
SAPR_TMA=0 timeout 120 python tools/vit_bench.py 100000 10 2>&1 | tail -1
timeout 120 python tools/vit_bench.py 100000 10 2>&1 | tail -1
timeout 120 python tools/vit_bench.py 94720 10 2>&1 | tail -1
timeout 300 python tools/estep_bench.py 200000 5 2>&1 | tail -2
