#!/usr/bin/env bash
set -x
timeout 900 python -m pytest tests/test_gpu_rd.py -m gpu -q -x -s -k "default_mode_trace" 2>&1 | tail -12
