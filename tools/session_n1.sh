#!/bin/bash
# one-GPU session: walk / build changes -- device-path tests, then the few-nodes threshold sweep of the tree build
cd /root/repo
timeout 1200 python -m pytest tests/test_gpu_device_tree.py tests/test_gpu_device_multirank.py tests/test_gpu_resident.py tests/test_gpu_midfield.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2x_pytest.txt; cat gpurun_out/r2x_pytest.txt
for few in 0 296 148 74; do
  echo "== P2P_B200_FEW_NODES=$few" >> gpurun_out/r2x_few.txt
  P2P_B200_FEW_NODES=$few timeout 200 python tools/device_step.py 256 32 3 --resident 2>&1 | tail -2 >> gpurun_out/r2x_few.txt
  P2P_B200_FEW_NODES=$few timeout 200 python tools/device_step.py 128 32 3 --resident 2>&1 | tail -1 >> gpurun_out/r2x_few.txt
done
cat gpurun_out/r2x_few.txt
