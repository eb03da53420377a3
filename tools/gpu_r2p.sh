#!/bin/bash
# round 2, call P: full GPU suite + default bench line
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -15 > gpurun_out/tests_p.log; tail -6 gpurun_out/tests_p.log
timeout 900 python bench.py > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; tail -c 600 gpurun_out/bench_p.json; tail -3 gpurun_out/bench_p.err
