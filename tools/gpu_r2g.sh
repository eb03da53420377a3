#!/bin/bash
# round 2, call G: ncu source-level capture of the grouped E-step
mkdir -p gpurun_out
timeout 300 python tools/prof_one.py > gpurun_out/prof_one_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_estep_grouped" -s 2 -c 1 -f -o gpurun_out/prof_eg python tools/prof_one.py > gpurun_out/ncu_eg.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_eg.log
