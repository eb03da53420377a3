#!/bin/bash
# round 2, call P: full GPU suite
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -15 > gpurun_out/tests_p.log; tail -6 gpurun_out/tests_p.log
