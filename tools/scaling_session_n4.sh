#!/bin/bash
# 4-GPU session: NCCL parity at 2 and 4 ranks, strong scaling of the whole step at 256^3 and 512^3, clustered 512^3 with relaxation
cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_nccl_multirank.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2q_pytest_nccl.txt; cat gpurun_out/r2q_pytest_nccl.txt
for W in 2 4; do timeout 300 $TR --nproc-per-node $W --master-port 2961$W tests/tools/nccl_parity.py gpurun_out/r2q_nccl_parity_n$W.json > /dev/null 2> gpurun_out/r2q_nccl_parity_n$W.err; echo rcparity$W $?; done
timeout 240 $TR --nproc-per-node 4 --master-port 29641 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2q_bench256_n4.json 2> gpurun_out/r2q_bench256_n4.err; echo rc256 $?
timeout 300 $TR --nproc-per-node 4 --master-port 29642 bench.py --gpus 4 --nside 512 --steps 3 --warmup 3 > gpurun_out/r2q_bench512_n4.json 2> gpurun_out/r2q_bench512_n4.err; echo rc512 $?
timeout 300 $TR --nproc-per-node 4 --master-port 29643 bench.py --gpus 4 --nside 512 --clustered --relax 4 --steps 2 --warmup 3 --no-e2e --no-launch-count > gpurun_out/r2q_bench512c_n4.json 2> gpurun_out/r2q_bench512c_n4.err; echo rc512c $?
for f in gpurun_out/r2q_*.err; do echo == $f; grep -v "^\[W\|^W1\|^$\|^\*\*\*\|OMP_NUM\|NCCL version" $f | tail -4; done
