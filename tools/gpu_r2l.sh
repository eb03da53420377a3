#!/bin/bash
# round 2, call L: k_viterbi_v4 what-if experiments (which role sets the pace)
for e in 32 16 24 1 2 4 6 7 15 31 32; do echo "exp $e"; SAPR_V_EXP=$e timeout 120 python tools/vit_bench.py 94720 10 2>&1 | tail -1 | cut -c30-130; done
