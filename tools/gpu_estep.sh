#!/bin/bash
# quick E-step iteration: parity tests that touch the E-step, then the bench legs, then the pipeline trace
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 -k "estep or fit or train or baum or scale or moderate or ragged or smoke or stats" 2>&1 | tail -15 > gpurun_out/tests_estep.log; tail -4 gpurun_out/tests_estep.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --ergodic-utts 0 --audio-utts 0 --no-cfg1 > gpurun_out/bench_estep.log 2>&1; python - <<'PY'
import json
try:
    l=json.loads(open('gpurun_out/bench_estep.log').read().strip().splitlines()[-1]); e=l['estep']
    print('viterbi ms', l['ms_per_step'], 'estep ms', e['ms_per_iteration'], 'fwdbwd', e['roofline']['fwdbwd_kernel_ms'], 'stats', e['roofline']['stats_kernel_ms'], 'frac', e['roofline']['frac'])
except Exception as ex:
    print('bench failed', ex); print(open('gpurun_out/bench_estep.log').read()[-2000:])
PY
timeout 200 python tools/et_trace.py > gpurun_out/et_trace_print.txt 2>&1; echo trace rc=$?
