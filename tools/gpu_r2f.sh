#!/bin/bash
# round 2, call F: grouped E-step skeleton without TMA loads; Viterbi with batched accumulator loads
mkdir -p gpurun_out
for e in 0 14 30 16; do SAPR_EG_EXP=$e timeout 200 python tools/estep_bench.py 200000 4 grouped 2>&1 | tail -1 | cut -c1-200; done
timeout 120 python tools/vit_bench.py 94720 10 2>&1 | tail -1
SAPR_TM_BATCH=1 timeout 120 python tools/vit_bench.py 94720 10 2>&1 | tail -1
