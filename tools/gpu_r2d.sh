#!/bin/bash
# round 2, call D: grouped E-step with the greedy MMA issuer; what-if experiments (recursion / conversion arithmetic removed)
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q -k "grouped" ) 2>&1 | tail -4
for e in 0 2 4 6; do SAPR_EG_EXP=$e timeout 200 python tools/estep_bench.py 200000 4 grouped 2>&1 | tail -1 | cut -c1-260; done
