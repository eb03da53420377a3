#!/bin/bash
mkdir -p gpurun_out
for pf in 1 2 4 6 10 100000; do
SAPR_ET_PF=$pf timeout 300 python bench.py --steps 2 --warmup 3 --utts 2560 --no-cpu --no-e2e --ergodic-utts 0 > gpurun_out/bench_pf.log 2>&1; python - <<PY
import json
l=json.loads(open('gpurun_out/bench_pf.log').read().strip().splitlines()[-1]); e=l['estep']
print('pf $pf estep ms', e['ms_per_iteration'], 'fwdbwd', e['roofline']['fwdbwd_kernel_ms'])
PY
done
