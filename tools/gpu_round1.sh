#!/usr/bin/env bash
# First GPU pass: parity tests, FP64 peak, short benches.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
./tools/bin/fp64_peak > gpurun_out/fp64_peak.json 2>&1; cat gpurun_out/fp64_peak.json
for tb in 1 2 4; do
  timeout 600 python bench.py --nx 8192 --ny 8192 --tb $tb --steps 5 --warmup 3 --substeps 32 --no-cpu-baseline 2>&1 | tail -2 > gpurun_out/bench_8192_tb$tb.json
  cat gpurun_out/bench_8192_tb$tb.json
done
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -2 > gpurun_out/bench_16384.json; cat gpurun_out/bench_16384.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 > gpurun_out/bench_ref_16384.json; cat gpurun_out/bench_ref_16384.json
