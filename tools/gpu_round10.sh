#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for w in 64 128 192; do
YH_RK_W=$w timeout 900 python bench.py --mode rk4lap4 --nx 8192 --ny 8192 --steps 3 --warmup 3 --substeps 8 --e2e-substeps 8 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rk4lap4 8192 W=$w', round(d['value'],2), 'Gcell/s')"
done
timeout 600 python tools/small_sweep.py 2>&1 | grep "rk4lap4" -A1 | tail -9
