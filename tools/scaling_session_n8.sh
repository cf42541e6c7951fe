#!/bin/bash
# one 8-GPU session: strong scaling of the whole step at 256^3, 512^3 and 1024^3
cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2y_bench256_n8.json 2> gpurun_out/r2y_bench256_n8.err; echo rc256 $?
timeout 300 $TR --master-port 29542 bench.py --gpus 8 --nside 512 --steps 3 --warmup 3 > gpurun_out/r2y_bench512_n8.json 2> gpurun_out/r2y_bench512_n8.err; echo rc512 $?
timeout 400 $TR --master-port 29543 bench.py --gpus 8 --nside 1024 --steps 2 --warmup 3 --no-launch-count > gpurun_out/r2y_bench1024_n8.json 2> gpurun_out/r2y_bench1024_n8.err; echo rc1024 $?
for f in gpurun_out/r2y_bench*_n8.err; do echo == $f; grep -v "^\[W\|^W1\|^$\|^\*\*\*\|OMP_NUM\|NCCL version\|_warn_once\|UserWarning" $f | tail -4; done
