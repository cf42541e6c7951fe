#!/bin/bash
# 8-GPU session: the clustered 1024^3 box with split relaxation
cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29544 bench.py --gpus 8 --nside 1024 --clustered --relax 4 --steps 2 --warmup 3 --no-e2e --no-launch-count > gpurun_out/r2p_bench1024c_n8.json 2> gpurun_out/r2p_bench1024c_n8.err; echo rc1024c $?
grep -v "^\[W\|^W1\|^$" gpurun_out/r2p_bench1024c_n8.err | tail -12
