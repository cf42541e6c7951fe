#!/bin/bash
# one-GPU session: the ncu launch list of the bench command (after the same command ran clean without ncu)
cd /root/repo
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-launch-count > gpurun_out/r2E_plain.json 2> gpurun_out/r2E_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2E_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-launch-count > gpurun_out/r2E_ncu.log 2>&1; echo rcncu $?
