#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for tb in 1 4; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:rd_euler_stream -s 6 -c 1 \
     -o gpurun_out/prof_tb$tb -f python bench.py --nx 8192 --ny 8192 --tb $tb --steps 1 --warmup 3 --substeps 8 --no-cpu-baseline > gpurun_out/ncu_tb$tb.log 2>&1
  tail -3 gpurun_out/ncu_tb$tb.log
done
