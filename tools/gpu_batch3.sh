#!/usr/bin/env bash
TAG=${1:-b3}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_march.py tests/test_gpu_aux.py -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.log
for pdl in 1 0; do
  for ar in exact fast; do
    YH_MARCH_PDL=$pdl YH_ARITH=$ar timeout 200 python tools/rk_probe.py 512 4000 rk4lap4 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
  done
done
YH_MARCH_R=3 YH_ARITH=fast timeout 200 python tools/rk_probe.py 512 4000 rk4lap4 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
YH_MARCH_R=4 YH_ARITH=exact timeout 200 python tools/rk_probe.py 512 4000 rk4lap4 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
YH_RD_PATH=tile timeout 200 python tools/rk_probe.py 1024 500 rk4lap4 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
YH_RD_PATH=stream timeout 200 python tools/rk_probe.py 1024 500 rk4lap4 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
YH_RD_PATH=tile timeout 200 python tools/rk_probe.py 1024 500 rk4 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
YH_RD_PATH=stream timeout 200 python tools/rk_probe.py 1024 500 rk4 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
for cfg in "8 28" "16 14" "8 40" "16 20" "4 56"; do
  set -- $cfg
  YH_SLAB_PIPE_CHUNKS=$1 YH_SLAB_PIPE_LEVELS=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-modes 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_bench_pipe.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench_pipe.json')); print('chunks $1 levels $2: value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))" | tee -a gpurun_out/${TAG}_pipe_tune.txt
done
