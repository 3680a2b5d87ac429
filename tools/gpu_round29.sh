#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_rd.py -m gpu -q -x -k "default_mode_trace" 2>&1 | tail -25; done
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
