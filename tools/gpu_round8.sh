#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 1200 python tools/config_compare.py 2>&1 | tail -6
