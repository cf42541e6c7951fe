#!/bin/bash
# one-GPU session: tree-build tests with the block-centric levels, then the resident step timing
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_device_tree.py tests/test_gpu_resident.py tests/test_gpu_midfield.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2C_pytest.txt; cat gpurun_out/r2C_pytest.txt
timeout 200 python tools/device_step.py 256 32 3 --resident 2>&1 | tail -1 > gpurun_out/r2C_devstep.txt
timeout 200 python tools/device_step.py 128 32 3 --resident 2>&1 | tail -1 >> gpurun_out/r2C_devstep.txt
timeout 200 python tools/device_step.py 128 32 2 --clustered 2>&1 | tail -1 >> gpurun_out/r2C_devstep.txt
cat gpurun_out/r2C_devstep.txt
