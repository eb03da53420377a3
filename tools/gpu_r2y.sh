#!/bin/bash
# round 2, call Y: source-level profiles of the small latency-bound kernels behind the main Viterbi launch
mkdir -p gpurun_out
for k in k_redo_f64 k_redo_emission k_viterbi_finish_fast; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$k -s 3 -c 1 -o gpurun_out/small_$k -f python tools/vit_bench.py 100000 2 > gpurun_out/small_$k.log 2>&1
  ncu -i gpurun_out/small_$k.ncu-rep --page source --csv > gpurun_out/small_${k}_src.csv 2>/dev/null
  echo "== $k"; ncu -i gpurun_out/small_$k.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; v=r[2]
for k in ('gpu__time_duration.sum','smsp__inst_executed.sum','launch__registers_per_thread'): print(k, v[h.index(k)])"
  python tools/ncu_src_top.py gpurun_out/small_${k}_src.csv 14
done
