#!/usr/bin/env bash
# pipelined yh_slab_run_host: parity (C++ host, N slabs on the visible GPUs), then the bench e2e leg A/B
TAG=${1:-p1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slab_driver.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_slab.log
for pipe in 1 0; do
  YH_SLAB_PIPE=$pipe timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-modes 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_bench_pipe${pipe}.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench_pipe${pipe}.json')); print('pipe=$pipe value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['impl_config']['checksum'])"
done
