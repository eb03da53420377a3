#!/bin/bash
# round 2, call C: grouped E-step parity diagnostics + ncu source-level capture of the two new kernels
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -q -k "grouped" ) 2>&1 | tail -60 > gpurun_out/tests_c.log; tail -5 gpurun_out/tests_c.log
timeout 300 python tools/prof_one.py > gpurun_out/prof_one_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_estep_grouped|k_viterbi_tma" -s 2 -c 2 -o gpurun_out/prof_r2c python tools/prof_one.py > gpurun_out/ncu_r2c.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_r2c.log
