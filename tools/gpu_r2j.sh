#!/bin/bash
# round 2, call J: ncu source-level capture of k_viterbi_v3
mkdir -p gpurun_out
timeout 300 python tools/prof_vit.py > gpurun_out/prof_vit_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_viterbi_v3" -s 1 -c 1 -f -o gpurun_out/prof_v3 python tools/prof_vit.py > gpurun_out/ncu_v3.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_v3.log
