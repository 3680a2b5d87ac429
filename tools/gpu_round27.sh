#!/usr/bin/env bash
mkdir -p gpurun_out
for w in 64 128 256; do for ry in 0 64 128 256 512; do
  if [ $ry = 0 ]; then unset YH_FAST_RY; else export YH_FAST_RY=$ry; fi
  v=$(YH_FAST_W=$w timeout 300 python bench.py --workload sweep --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; print(round(json.loads(sys.stdin.read())['value'],1))")
  echo "W=$w RY=$ry : $v"
done; done | tee gpurun_out/sweep_tune.txt
