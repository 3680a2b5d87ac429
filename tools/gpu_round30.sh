#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
for w in 64 128 256; do
  v=$(YH_FAST_W=$w timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-substeps 64 2>/dev/null | tail -1 | python -c "import json,sys; print(round(json.loads(sys.stdin.read())['value'],1))")
  echo "16384^2 T=4 W=$w : $v"
done | tee gpurun_out/w_sweep_16384.txt
