#!/bin/bash
# ncu --set full of the launch pair at the headline batch (full rounds: k_viterbi_v4<5,0,0,0>; partial round: <5,0,0,1>, two CTAs per tile)
mkdir -p gpurun_out
CMD="python tools/vit_bench.py 100000 2"
$CMD > gpurun_out/plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_viterbi_v4|k_redo_fused|k_viterbi_finish_v3" -s 15 -c 5 -f -o gpurun_out/prof_pair $CMD > gpurun_out/ncu_pair.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/plain4.log | cut -c1-200
