#!/usr/bin/env bash
TAG=${1:-b7}
mkdir -p gpurun_out
for a in "256 256 2 60 sr" "256 256 3 40 sr" "256 384 3 40 sr" "512 512 2 200 sr"; do
  echo "== $a" | tee -a gpurun_out/${TAG}_sr_driver.txt
  YH_SR_DEBUG=1 timeout 40 yolohtli_b200/lib/yh_slab_driver $a 1 > gpurun_out/${TAG}_out.txt 2> gpurun_out/${TAG}_err.txt
  echo "rc=$?" | tee -a gpurun_out/${TAG}_sr_driver.txt
  tail -2 gpurun_out/${TAG}_out.txt | tee -a gpurun_out/${TAG}_sr_driver.txt
  (head -3 gpurun_out/${TAG}_err.txt; echo ...; tail -3 gpurun_out/${TAG}_err.txt) | tee -a gpurun_out/${TAG}_sr_driver.txt
done
