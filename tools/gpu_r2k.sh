#!/bin/bash
# round 2, call K: k_viterbi_v5 (three accumulator stages) parity + timing against v4
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_scale.py -m gpu -q -k "viterbi or near_ties" ) 2>&1 | tail -40 > gpurun_out/tests_k.log; tail -3 gpurun_out/tests_k.log
SAPR_VK=4 timeout 120 python tools/vit_bench.py 100000 10 2>&1 | tail -1 | cut -c1-150
timeout 120 python tools/vit_bench.py 100000 10 2>&1 | tail -1 | cut -c1-150
SAPR_VK=4 timeout 120 python tools/vit_bench.py 94720 10 2>&1 | tail -1 | cut -c1-150
timeout 120 python tools/vit_bench.py 94720 10 2>&1 | tail -1 | cut -c1-150
