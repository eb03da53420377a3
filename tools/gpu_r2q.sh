#!/bin/bash
# round 2, call Q: word-exactness pass -- parity tests, cost
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_scale.py tests/test_gpu_train_eval.py -m gpu -q -x -k "viterbi or near_ties or train_eval" ) 2>&1 | tail -30 > gpurun_out/tests_q.log; tail -4 gpurun_out/tests_q.log
SAPR_EXACT_WORDS=0 timeout 120 python tools/vit_bench.py 100000 20 2>&1 | tail -1 | cut -c1-230
timeout 120 python tools/vit_bench.py 100000 20 2>&1 | tail -1 | cut -c1-230
timeout 120 python tools/vit_bench.py 12500 20 2>&1 | tail -1 | cut -c1-230
