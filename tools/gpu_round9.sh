#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for tb in 1 2 4; do
  timeout 600 python bench.py --nx 8192 --ny 8192 --tb $tb --steps 5 --warmup 3 --substeps 32 --e2e-substeps 32 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('8192 tb=$tb', round(d['value'],1), 'Gcell/s frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))"
done
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('16384 tb=4', round(d['value'],1), 'Gcell/s frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))"
