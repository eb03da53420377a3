#!/bin/bash
# round 2, call X: branch-free back-trace kernels: parity suite, then step / finish / redo times
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for n in 100000 94720 5376; do
  timeout 120 python tools/vit_bench.py $n 20 2>&1 | tail -1 | cut -c1-330
done
