#!/bin/bash
# launch list of the bench at the headline Viterbi workload (100 k utterances: full rounds + partial round), smaller side legs
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --estep-utts 37888 --ergodic-utts 18944 --no-cpu --no-e2e"
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches100k.csv $CMD > gpurun_out/ncu_list3.log 2>&1
echo "ncu list rc=$?"; tail -1 gpurun_out/plain3.log | cut -c1-200
