#!/bin/bash
# round 2, call S (8 GPUs): comm check + bench line at 8 GPUs (weak + strong cfg 2, 1 M-utterance E-step with the native all-reduce)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/comm_check.py 2>&1 | grep "comm_check"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --ergodic-utts 0 --audio-utts 0 --no-cfg1 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; tail -c 400 gpurun_out/bench_8gpu.json; tail -2 gpurun_out/bench_8gpu.err
