#!/bin/bash
# one-GPU session: parity at scale, the default bench line, 512^3
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity_scale.py tests/test_gpu_device_multirank.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2D_pytest.txt; cat gpurun_out/r2D_pytest.txt
timeout 400 python bench.py > gpurun_out/r2D_bench256_n1.json 2> gpurun_out/r2D_bench256_n1.err; echo rc256 $?
timeout 300 python bench.py --nside 512 --steps 3 --warmup 3 --no-e2e > gpurun_out/r2D_bench512_n1.json 2> gpurun_out/r2D_bench512_n1.err; echo rc512 $?
