#!/bin/bash
# round 2, call H (re-entry): full GPU suite, then the round-2 kernels timed against the round-1 ones
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -25 > gpurun_out/tests_h.log; tail -8 gpurun_out/tests_h.log
SAPR_TMA=0 timeout 120 python tools/vit_bench.py 100000 10 2>&1 | tail -1
timeout 120 python tools/vit_bench.py 100000 10 2>&1 | tail -1
timeout 120 python tools/vit_bench.py 94720 10 2>&1 | tail -1
timeout 300 python tools/estep_bench.py 200000 5 2>&1 | tail -1
