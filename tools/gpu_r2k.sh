#!/bin/bash
# round 2, call K: k_viterbi_v4 (role-split) parity + timing against v3, pipeline timeline
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_scale.py -m gpu -q -k "viterbi" ) 2>&1 | tail -40 > gpurun_out/tests_k.log; tail -3 gpurun_out/tests_k.log
SAPR_VK=3 timeout 120 python tools/vit_bench.py 100000 10 2>&1 | tail -1 | cut -c1-150
timeout 120 python tools/vit_bench.py 100000 10 2>&1 | tail -1 | cut -c1-150
timeout 120 python tools/vit_bench.py 94720 10 2>&1 | tail -1 | cut -c1-150
timeout 200 python tools/v4_trace.py gpurun_out/v4_trace.txt 100 24 > gpurun_out/v4_trace.log 2>&1; tail -1 gpurun_out/v4_trace.log
