#!/bin/bash
# tests + launch list + one full ncu capture of the dominant kernel (run under gpurun)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -60 > gpurun_out/tests.log; tail -15 gpurun_out/tests.log
CMD="python bench.py --steps 2 --warmup 3 --utts 20000 --estep-utts 20000 --no-cpu --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_viterbi_fused -s 3 -c 1 -o gpurun_out/prof_viterbi $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_estep_fused|k_stats_diag" -s 4 -c 2 -o gpurun_out/prof_estep $CMD > gpurun_out/ncu_full2.log 2>&1
echo "ncu full2 rc=$?"
tail -3 gpurun_out/ncu_full.log
