#!/bin/bash
# round 2, call M: k_viterbi_v4 pipeline timeline
mkdir -p gpurun_out
timeout 200 python tools/v4_trace.py gpurun_out/v4_trace.txt 100 24 > gpurun_out/v4_trace.log 2>&1; tail -3 gpurun_out/v4_trace.log
timeout 120 python tools/vit_bench.py 94720 10 2>&1 | tail -1 | cut -c1-150
