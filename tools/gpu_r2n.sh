#!/bin/bash
# round 2, call N: grouped E-step parity + timing
( timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q -k "grouped" ) 2>&1 | tail -4
timeout 300 python tools/estep_bench.py 200000 5 grouped 2>&1 | tail -1 | cut -c1-330
