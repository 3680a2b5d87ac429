#!/usr/bin/env bash
# Small-sheet default mode: marching tile kernel (rd_tile_march.cu) vs the first tile kernel; parity first.
#   gpurun --timeout 900 -- 'bash tools/gpu_march.sh TAG'
mkdir -p gpurun_out
TAG=${1:-m1}
timeout 600 python -m pytest tests/test_gpu_march.py tests/test_gpu_slab_driver.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_march.log
for r in 6 4; do
  for ar in exact fast; do
    for args in "512 2000 rk4lap4" "512 2000 rk4" "256 2000 rk4lap4" "768 1000 rk4lap4"; do
      YH_MARCH_R=$r YH_ARITH=$ar timeout 200 python tools/rk_probe.py $args 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
    done
  done
done
NCU="ncu --set full --clock-control none --import-source on -f"
export YH_GRAPHS=0
cap() { # name kernel-regex skip probe-args...
  local name=$1 rex=$2 skip=$3; shift 3
  timeout 300 $NCU -k regex:$rex -s $skip -c 1 -o gpurun_out/${TAG}_$name python tools/rk_probe.py "$@" 2>&1 | grep rk_probe
  ncu -i gpurun_out/${TAG}_$name.ncu-rep --page raw --csv > gpurun_out/${TAG}_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_$name.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_${name}_source.csv 2>/dev/null
  python tools/ncu_mix.py gpurun_out/${TAG}_${name}_source.csv 24 > gpurun_out/${TAG}_${name}_mix.txt 2>/dev/null
  rm -f gpurun_out/${TAG}_$name.ncu-rep
}
cap tile_march_512 rd_tile_march 8 512 16 rk4lap4
YH_ARITH=fast cap tile_march_512_fast rd_tile_march 8 512 16 rk4lap4
