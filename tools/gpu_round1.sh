#!/bin/bash
# tests + full bench + launch list + one full ncu capture of the dominant kernel (run under gpurun)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -60 > gpurun_out/tests.log; tail -5 gpurun_out/tests.log
python bench.py > gpurun_out/bench_full.log 2>&1; tail -1 gpurun_out/bench_full.log | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-300
CMD="python bench.py --steps 2 --warmup 3 --utts 18944 --estep-utts 18944 --no-cpu --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_viterbi_tc -s 3 -c 1 -o gpurun_out/prof_viterbi_tc $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
