#!/bin/bash
# round 2, call G: ncu source-level capture of the grouped E-step (TMEM A operand, greedy MMA issuer)
mkdir -p gpurun_out
timeout 300 python tools/prof_one.py > gpurun_out/prof_one_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_estep_grouped" -s 2 -c 1 -o gpurun_out/prof_r2g python tools/prof_one.py > gpurun_out/ncu_r2g.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_r2g.log
