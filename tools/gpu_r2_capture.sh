#!/usr/bin/env bash
# Round-2 evidence for the kernels bench.py actually times: one `ncu --set full` capture each
# (source page included), the launch list of the default bench command, the FP64 issue peak.
#   gpurun --timeout 900 -- 'bash tools/gpu_r2_capture.sh <tag>'
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
tools/bin/fp64_peak | tee gpurun_out/${TAG}_fp64_peak.txt
NCU="ncu --set full --clock-control none --import-source on -f"
export YH_GRAPHS=0
timeout 300 $NCU -k regex:rd_euler_stream -s 2 -c 1 -o gpurun_out/${TAG}_euler_tb4_16384 python tools/rk_probe.py 16384 8 euler 2>&1 | grep rk_probe
timeout 200 $NCU -k regex:rd_euler_stream -s 2 -c 1 -o gpurun_out/${TAG}_euler_tb1_8192 python tools/rk_probe.py 8192 8 euler1 2>&1 | grep rk_probe
timeout 200 $NCU -k regex:rd_rk_stream -s 4 -c 1 -o gpurun_out/${TAG}_rk4lap4_8192 python tools/rk_probe.py 8192 4 rk4lap4 2>&1 | grep rk_probe
timeout 120 $NCU -k regex:rd_tile_rk -s 8 -c 1 -o gpurun_out/${TAG}_tile_rk_512 python tools/rk_probe.py 512 16 rk4lap4 2>&1 | grep rk_probe
timeout 120 $NCU -k regex:rd_tile_euler -s 8 -c 1 -o gpurun_out/${TAG}_tile_euler_512 python tools/rk_probe.py 512 64 euler 2>&1 | grep rk_probe
unset YH_GRAPHS
# plain timings of the same commands (never read a number from a run under ncu)
for args in "16384 16 euler" "8192 16 euler1" "8192 8 rk4lap4" "512 512 rk4lap4" "512 2048 euler"; do
  timeout 200 python tools/rk_probe.py $args 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
done
