#!/usr/bin/env bash
mkdir -p gpurun_out
for w in 64 128 192; do
  v=$(YH_RK_W=$w timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mode rk4lap4 --nx 8192 --ny 8192 --substeps 16 --e2e-substeps 16 2>/dev/null | tail -1 | python -c "import json,sys; print(round(json.loads(sys.stdin.read())['value'],2))")
  echo "8192^2 RK4+lap4 W=$w : $v"
done | tee gpurun_out/w_sweep_rk_8192.txt
for ry in 0 256 512 1024 2048; do
  if [ $ry = 0 ]; then unset YH_FAST_RY; else export YH_FAST_RY=$ry; fi
  v=$(timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-substeps 64 2>/dev/null | tail -1 | python -c "import json,sys; print(round(json.loads(sys.stdin.read())['value'],1))")
  echo "16384^2 T=4 W=128 RY=$ry : $v"
done | tee -a gpurun_out/w_sweep_16384.txt
