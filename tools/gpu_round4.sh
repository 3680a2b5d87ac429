#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 python -m tests.golden.make_golden > gpurun_out/golden.log 2>&1; tail -25 gpurun_out/golden.log
mkdir -p gpurun_out/golden && cp tests/golden/*.npz gpurun_out/golden/
for w in 256 128; do
  YH_FAST_W=$w timeout 600 ncu --set full --clock-control none --import-source on -k regex:rd_euler_stream -s 6 -c 1 \
     -o gpurun_out/prof2_w${w}_tb4 -f python bench.py --nx 8192 --ny 8192 --tb 4 --steps 1 --warmup 3 --substeps 8 --no-cpu-baseline > gpurun_out/ncu2_w$w.log 2>&1
done
