#!/usr/bin/env bash
# Euler kernel A/B: column pairs per thread (rd_fast.cu) vs quads (rd_quad.cu); parity first.
mkdir -p gpurun_out
TAG=${1:-q1}
timeout 900 python -m pytest tests/test_gpu_rd.py -m gpu -q -x -k "quad" 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_quad.log
for kern in pair quad; do
  for args in "16384 16 euler" "8192 16 euler2" "8192 16 euler1"; do
    YH_EULER_KERNEL=$kern timeout 200 python tools/rk_probe.py $args 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
  done
done
YH_EULER_KERNEL=quad YH_FAST_W=256 timeout 200 python tools/rk_probe.py 16384 16 euler 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
YH_EULER_KERNEL=quad YH_FAST_W=256 timeout 200 python tools/rk_probe.py 8192 16 euler1 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
