#!/bin/bash
# one 8-GPU session: strong scaling of the whole step at 256^3, 512^3, 1024^3 and the clustered 1024^3 box
cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
free -g | head -2 > gpurun_out/r2n_mem.txt; nproc >> gpurun_out/r2n_mem.txt
AVAIL=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
E2E=""; if [ "$AVAIL" -lt 300 ]; then E2E="--no-e2e"; fi
timeout 240 $TR --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2n_bench256_n8.json 2> gpurun_out/r2n_bench256_n8.err; echo rc256 $?
timeout 300 $TR --master-port 29542 bench.py --gpus 8 --nside 512 --steps 3 --warmup 3 > gpurun_out/r2n_bench512_n8.json 2> gpurun_out/r2n_bench512_n8.err; echo rc512 $?
timeout 400 $TR --master-port 29543 bench.py --gpus 8 --nside 1024 --steps 2 --warmup 3 $E2E --no-launch-count > gpurun_out/r2n_bench1024_n8.json 2> gpurun_out/r2n_bench1024_n8.err; echo rc1024 $?
timeout 500 $TR --master-port 29544 bench.py --gpus 8 --nside 1024 --clustered --relax 4 --steps 2 --warmup 3 --no-e2e --no-launch-count > gpurun_out/r2n_bench1024c_n8.json 2> gpurun_out/r2n_bench1024c_n8.err; echo rc1024c $?
for f in gpurun_out/r2n_bench*_n8.err; do echo == $f; tail -4 $f; done
