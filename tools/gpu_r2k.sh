#!/bin/bash
# round 2, call K: k_viterbi_v4 variants parity + timing
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_scale.py -m gpu -q -k "viterbi or near_ties" ) 2>&1 | tail -40 > gpurun_out/tests_k.log; tail -3 gpurun_out/tests_k.log
for i in 1 2; do SAPR_EXACT_WORDS=0 timeout 120 python tools/vit_bench.py 100000 20 2>&1 | tail -1 | cut -c1-150; done
SAPR_EXACT_WORDS=0 timeout 120 python tools/vit_bench.py 94720 20 2>&1 | tail -1 | cut -c1-150
