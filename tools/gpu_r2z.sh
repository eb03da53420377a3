#!/bin/bash
# round 2, call Z: serialised per-kernel durations of one 100 k Viterbi step (ncu launch list)
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/vit_launches.csv python tools/vit_bench.py 100000 3 > gpurun_out/vit_launches.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/vit_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); gi = hdr.index("Grid Size")
for r in rows[-14:]:
    print(r[ki][:40], r[gi], r[vi])
PY
