#!/bin/bash
# round 2, call L: k_viterbi_v4 what-if experiments (which role sets the pace)
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv,noheader
for e in 0 15 63 14 62 6 7 0 15 63; do echo "exp $e"; SAPR_V_EXP=$e timeout 120 python tools/vit_bench.py 94720 10 2>&1 | tail -1 | cut -c30-130; done
