#!/bin/bash
# full default bench line at N GPUs of one box: gpurun --gpus N -- bash tools/gpu_bench_n.sh N
N=${1:-4}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; tail -c 200 gpurun_out/bench_${N}gpu.json; tail -2 gpurun_out/bench_${N}gpu.err | cut -c1-200
