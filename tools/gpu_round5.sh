#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_rd.py -m gpu -q -x -k "temporal or slab" 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for w in 256 128; do for tb in 1 2 4; do
  YH_FAST_W=$w timeout 600 python bench.py --nx 8192 --ny 8192 --tb $tb --steps 5 --warmup 3 --substeps 32 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('W=$w tb=$tb', round(d['value'],1), 'Gcell/s frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))"
done; done
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('16384 tb=4', round(d['value'],1), 'Gcell/s frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))"
YH_FAST_W=128 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('16384 W128 tb=4', round(d['value'],1), 'Gcell/s frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))"
