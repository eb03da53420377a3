#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q -k "viterbi or near_ties" 2>&1 | tail -3
for hf in 0 1; do
  for n in 100000 120000; do
    SAPR_V_HALF=$hf timeout 120 python tools/vit_bench.py $n 20 2>&1 | tail -1 | cut -c1-330
  done
done
