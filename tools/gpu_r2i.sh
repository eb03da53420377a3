#!/bin/bash
# round 2, call I: k_viterbi_v3 parity + timing; traceback of the pickle test
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_custom_hmm.py -m gpu -x -q -k pickle_decodes 2>&1 | tail -40 > gpurun_out/tests_i_pickle.log
( timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_scale.py -m gpu -q -k "viterbi" ) 2>&1 | tail -40 > gpurun_out/tests_i.log; tail -15 gpurun_out/tests_i.log
SAPR_V3=0 timeout 120 python tools/vit_bench.py 100000 10 2>&1 | tail -1
timeout 120 python tools/vit_bench.py 100000 10 2>&1 | tail -1
timeout 120 python tools/vit_bench.py 94720 10 2>&1 | tail -1
