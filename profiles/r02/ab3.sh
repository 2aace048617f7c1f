#!/bin/bash
# round 2, GPU call 3 (2 GPUs): peer-pull multi-GPU path -- parity tests, then bench at N = 2 (default, and with TAD_PEER_PULL=0)
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/ab3_topo.txt 2>&1
timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -x -q -s > gpurun_out/ab3_tests.log 2>&1; echo "rc=$?" >> gpurun_out/ab3_tests.log
tail -25 gpurun_out/ab3_tests.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2"
timeout 300 $T --steps 10 --warmup 3 --no-cpu > gpurun_out/ab3_n2.json 2> gpurun_out/ab3_n2.err; tail -c 2500 gpurun_out/ab3_n2.json; tail -5 gpurun_out/ab3_n2.err
TAD_PEER_PULL=0 timeout 300 $T --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab3_n2_nccl.json 2> gpurun_out/ab3_n2_nccl.err; tail -c 800 gpurun_out/ab3_n2_nccl.json
TAD_DETECT_MODE=1 timeout 300 $T --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab3_n2_direct.json 2> gpurun_out/ab3_n2_direct.err; tail -c 800 gpurun_out/ab3_n2_direct.json
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab3_n1.json 2> gpurun_out/ab3_n1.err; tail -c 600 gpurun_out/ab3_n1.json
