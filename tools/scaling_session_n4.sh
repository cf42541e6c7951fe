#!/bin/bash
# 4-GPU session: strong scaling of the whole step at 256^3 and 512^3 (NCCL parity at 2 and 4 ranks: tests/test_gpu_nccl_multirank.py)
cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 4 --master-port 29641 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2z_bench256_n4.json 2> gpurun_out/r2z_bench256_n4.err; echo rc256 $?
timeout 300 $TR --nproc-per-node 4 --master-port 29642 bench.py --gpus 4 --nside 512 --steps 3 --warmup 3 > gpurun_out/r2z_bench512_n4.json 2> gpurun_out/r2z_bench512_n4.err; echo rc512 $?
timeout 240 $TR --nproc-per-node 2 --master-port 29643 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2z_bench256_n2.json 2> gpurun_out/r2z_bench256_n2.err; echo rc256n2 $?
timeout 600 python -m pytest tests/test_gpu_nccl_multirank.py -m gpu -x -q 2>&1 | tail -3
