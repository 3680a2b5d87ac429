#!/usr/bin/env bash
TAG=${1:-b4}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_march.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.log
for kern in march stream; do
  for args in "1024 500 rk4holes" "512 1000 rk4holes" "2048 200 rk4holes" "4096 60 rk4holes"; do
    YH_SOLID_RK=$kern timeout 200 python tools/rk_probe.py $args 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
  done
done
