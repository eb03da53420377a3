#!/bin/bash
# round 2, call V: pipeline timelines of the single-product and the column-half Viterbi kernels
mkdir -p gpurun_out
SAPR_V_SPLIT=0 timeout 120 python tools/v4_trace.py gpurun_out/v4_trace_s0.txt 100 24 > gpurun_out/v4_trace_s0.log 2>&1
SAPR_V_SPLIT=1 timeout 120 python tools/v4_trace.py gpurun_out/v4_trace_s1.txt 100 24 > gpurun_out/v4_trace_s1.log 2>&1
python tools/v4_trace_stats.py gpurun_out/v4_trace_s0.txt 100 180
python tools/v4_trace_stats.py gpurun_out/v4_trace_s1.txt 100 180
