#!/bin/bash
# round 2, call E: grouped E-step: L2 prefetch distance sweep and what-if experiments
mkdir -p gpurun_out
for pf in 0 6 12 24; do SAPR_EG_PFD=$pf timeout 200 python tools/estep_bench.py 200000 4 grouped 2>&1 | tail -1 | cut -c1-200; done
for e in 6 14 2 4; do SAPR_EG_EXP=$e timeout 200 python tools/estep_bench.py 200000 4 grouped 2>&1 | tail -1 | cut -c1-200; done
SAPR_EG_PFD=0 SAPR_EG_EXP=14 timeout 200 python tools/estep_bench.py 200000 4 grouped 2>&1 | tail -1 | cut -c1-200
