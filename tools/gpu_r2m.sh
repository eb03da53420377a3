#!/bin/bash
# round 2, call M: k_viterbi_v4 pipeline timeline of the bare skeleton (no MMAs, no arithmetic, no stores)
mkdir -p gpurun_out
SAPR_V_EXP=15 timeout 200 python tools/v4_trace.py gpurun_out/v4_trace15.txt 100 16 > gpurun_out/v4_trace15.log 2>&1; grep "SM clock" gpurun_out/v4_trace15.log
SAPR_V_EXP=14 timeout 200 python tools/v4_trace.py gpurun_out/v4_trace14.txt 100 16 > gpurun_out/v4_trace14.log 2>&1; grep "SM clock" gpurun_out/v4_trace14.log
