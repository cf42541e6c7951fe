#!/bin/bash
# one-GPU session: smoke(), the default bench line, then the ncu launch list of the same bench command
cd /root/repo
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 > gpurun_out/r2A_smoke.txt; cat gpurun_out/r2A_smoke.txt
timeout 400 python bench.py > gpurun_out/r2A_bench256_n1.json 2> gpurun_out/r2A_bench256_n1.err; echo rc256 $?
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-launch-count > gpurun_out/r2A_plain.json 2> gpurun_out/r2A_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2A_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-launch-count > gpurun_out/r2A_ncu.log 2>&1; echo rcncu $?
