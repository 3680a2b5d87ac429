#!/usr/bin/env bash
TAG=${1:-b5}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_march.py tests/test_gpu_rd.py -m gpu -q -x -k "march or tile" 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.log
YH_ARITH=exact timeout 100 python tools/rk_probe.py 512 4000 rk4lap4 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
YH_ARITH=fast timeout 100 python tools/rk_probe.py 512 4000 rk4lap4 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
YH_ARITH=fast YH_MARCH_R=6 timeout 100 python tools/rk_probe.py 512 4000 rk4lap4 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
YH_ARITH=exact YH_MARCH_R=4 timeout 100 python tools/rk_probe.py 512 4000 rk4lap4 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
YH_ARITH=exact timeout 100 python tools/rk_probe.py 512 4000 rk4 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
for kern in march stream; do
  for args in "1024 500 rk4holes" "512 1000 rk4holes"; do
    YH_SOLID_RK=$kern timeout 100 python tools/rk_probe.py $args 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
  done
done
