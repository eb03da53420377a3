#!/bin/bash
# round 2, call K: k_viterbi_v4 (role-split) parity + timing against v3
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_scale.py -m gpu -q -k "viterbi" ) 2>&1 | tail -40 > gpurun_out/tests_k.log; tail -5 gpurun_out/tests_k.log
SAPR_VK=3 timeout 120 python tools/vit_bench.py 100000 10 2>&1 | tail -1
timeout 120 python tools/vit_bench.py 100000 10 2>&1 | tail -1
timeout 120 python tools/vit_bench.py 94720 10 2>&1 | tail -1
