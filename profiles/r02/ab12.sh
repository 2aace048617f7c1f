#!/bin/bash
# round 2, GPU call 12 (2 GPUs): last sanity check of the final build at N = 2 (symmetric regrow decision)
set -u
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab12_n2.json 2> gpurun_out/ab12_n2.err
tail -c 700 gpurun_out/ab12_n2.json; tail -2 gpurun_out/ab12_n2.err | cut -c1-200
