#!/bin/bash
# round 2, call O (2 GPUs): the C-ABI collective against torch.distributed, then the 2-GPU bench line (strong entry, native all-reduce)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/comm_check.py 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --ergodic-utts 0 --audio-utts 0 --no-cfg1 --no-e2e > gpurun_out/bench_2gpu_o.json 2> gpurun_out/bench_2gpu_o.err; tail -c 1500 gpurun_out/bench_2gpu_o.json; tail -3 gpurun_out/bench_2gpu_o.err
