#!/bin/bash
# one-GPU check session: device-path tests, kernel timing, the default bench line, then the ncu launch list of the same bench command
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_device_tree.py tests/test_gpu_parity.py tests/test_gpu_resident.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2v_pytest.txt; cat gpurun_out/r2v_pytest.txt
timeout 200 python tools/device_step.py 256 32 2 2>&1 | tail -2 > gpurun_out/r2v_devstep256.txt; cat gpurun_out/r2v_devstep256.txt
timeout 200 python tools/device_step.py 128 32 2 --clustered 2>&1 | tail -1 >> gpurun_out/r2v_devstep256.txt; tail -1 gpurun_out/r2v_devstep256.txt
timeout 400 python bench.py > gpurun_out/r2v_bench256_n1.json 2> gpurun_out/r2v_bench256_n1.err; echo rc256 $?
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-launch-count > gpurun_out/r2v_plain.json 2> gpurun_out/r2v_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2v_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-launch-count > gpurun_out/r2v_ncu.log 2>&1; echo rcncu $?
