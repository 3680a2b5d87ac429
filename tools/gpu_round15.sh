#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_rd.py tests/test_gpu_aux.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python tools/small_profile.py 2>&1 | grep "^nx" | grep -v holes | tee gpurun_out/small_profile.txt
