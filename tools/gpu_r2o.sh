#!/bin/bash
# round 2, call O (2 GPUs): reference arm under torchrun (core count), comm check, 2-GPU bench line
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29510 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-900 > gpurun_out/bench_ref_2gpu.json; python -c "
import json; l=json.loads(open('gpurun_out/bench_ref_2gpu.json').read()); print('ref arm under torchrun: cores', l['cpu_baseline']['cores'], 'value', l['value'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/comm_check.py 2>&1 | grep comm_check
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 300 gpurun_out/bench_2gpu.json
