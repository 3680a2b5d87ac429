#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rd_rk_stream -s 4 -c 1 -o gpurun_out/r1_rd_rk_stream_8192 -f \
   python bench.py --mode rk4lap4 --nx 8192 --ny 8192 --steps 1 --warmup 3 --substeps 2 --e2e-substeps 2 --no-cpu-baseline > gpurun_out/r1_rk_prof.log 2>&1
tail -2 gpurun_out/r1_rk_prof.log
