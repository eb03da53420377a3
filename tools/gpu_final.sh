#!/bin/bash
# round-end evidence (run under gpurun): GPU tests, smoke, both bench arms, launch list, full ncu captures of the dominant kernels
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --timeout 900 ) 2>&1 | tail -40 > gpurun_out/tests.log; tail -6 gpurun_out/tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-200
python bench.py > gpurun_out/bench_full.log 2>&1; tail -1 gpurun_out/bench_full.log | cut -c1-300
CMD="python bench.py --steps 2 --warmup 3 --utts 18944 --estep-utts 37888 --ergodic-utts 18944 --no-cpu --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
CMD2="python bench.py --steps 2 --warmup 3 --utts 18944 --estep-utts 37888 --ergodic-utts 0 --audio-utts 0 --no-cfg1 --no-cpu --no-e2e"
$CMD2 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_viterbi_v4|k_estep_tc|k_stats_diag8|k_viterbi_finish_v3" -s 6 -c 6 -f -o gpurun_out/prof_final $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
