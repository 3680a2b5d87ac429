#!/usr/bin/env bash
TAG=${1:-b6}
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_march.py tests/test_gpu_rd.py -m gpu -q -x -k "march or holes or masked or every_mode" 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.log
timeout 150 python -m pytest tests/test_gpu_slab_driver.py -m gpu -q -x -k "symmetry" 2>&1 | tail -12 | tee gpurun_out/${TAG}_pytest_sr.log
for a in "256 256 2 60 sr" "512 512 2 200 sr" "512 512 1 200 sr"; do timeout 60 yolohtli_b200/lib/yh_slab_driver $a 2 | tee -a gpurun_out/${TAG}_sr_driver.txt; done
YH_ARITH=exact timeout 60 python tools/rk_probe.py 512 4000 rk4lap4 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
YH_ARITH=fast timeout 60 python tools/rk_probe.py 512 4000 rk4lap4 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
for args in "1024 500 rk4holes" "512 1000 rk4holes" "2048 100 rk4holes"; do
  timeout 60 python tools/rk_probe.py $args 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
done
