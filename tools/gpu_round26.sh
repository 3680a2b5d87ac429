#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload sweep --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_sweep_n1.json | cut -c1-400
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_default.json | cut -c1-300
