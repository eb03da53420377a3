#!/bin/bash
# round 2, call J: ncu source-level capture of the cfg-2 Viterbi kernel
mkdir -p gpurun_out
timeout 300 python tools/prof_vit.py > gpurun_out/prof_vit_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_viterbi_v" -s 1 -c 1 -f -o gpurun_out/prof_v4 python tools/prof_vit.py > gpurun_out/ncu_v4.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_v4.log
