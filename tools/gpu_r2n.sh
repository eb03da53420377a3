#!/bin/bash
# round 2, call N: grouped E-step what-ifs (1 = stats right behind emission, 2 = no recursion arithmetic, 4 = no conversion arithmetic, 8 = no statistics MMAs, 16 = no TMA loads)
for e in 0 2 4 8 16 6 14 30; do echo "exp $e"; SAPR_EG_EXP=$e timeout 200 python tools/estep_bench.py 200000 4 grouped 2>&1 | tail -1 | cut -c1-230; done
