#!/usr/bin/env bash
# Round-2 evidence for the kernels bench.py times: one `ncu --set full` capture each (source page included),
# the launch list of the default bench command, the FP64 issue peak and the issue-slot microbenchmark.
#   gpurun --timeout 1200 -- 'bash tools/gpu_r2_capture.sh <tag>'
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
tools/bin/fp64_peak | tee gpurun_out/${TAG}_fp64_peak.txt
tools/bin/issue_mix > gpurun_out/${TAG}_issue_mix.txt
NCU="ncu --set full --clock-control none --import-source on -f"
export YH_GRAPHS=0
cap() { # name kernel-regex skip probe-args...
  local name=$1 rex=$2 skip=$3; shift 3
  timeout 300 $NCU -k regex:$rex -s $skip -c 1 -o gpurun_out/${TAG}_$name python tools/rk_probe.py "$@" 2>&1 | grep rk_probe
  # gpurun brings back at most 64 MiB: keep the exported pages, not the 15 MB report
  ncu -i gpurun_out/${TAG}_$name.ncu-rep --page raw --csv > gpurun_out/${TAG}_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_$name.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_${name}_source.csv 2>/dev/null
  python tools/ncu_mix.py gpurun_out/${TAG}_${name}_source.csv 24 > gpurun_out/${TAG}_${name}_mix.txt 2>/dev/null
  [ "$KEEP_REP" = "$name" ] || rm -f gpurun_out/${TAG}_$name.ncu-rep
}
KEEP_REP=euler_quad_tb4_16384
cap euler_quad_tb4_16384 rd_euler_quad 2 16384 8 euler
YH_ARITH=fast cap euler_quad_tb4_16384_fast rd_euler_quad 2 16384 8 euler
cap euler_stream_tb1_8192 rd_euler_stream 2 8192 8 euler1
cap rk_quad_lap4_8192 rd_rk_quad 4 8192 4 rk4lap4
YH_ARITH=fast cap rk_quad_lap4_8192_fast rd_rk_quad 4 8192 4 rk4lap4
cap tile_rk_512 rd_tile_rk 8 512 16 rk4lap4
cap tile_euler_512 rd_tile_euler 8 512 64 euler
unset YH_GRAPHS
# launch list of the default bench command (every kernel in the timed region, cold-cache serialised times)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 96 --csv --log-file gpurun_out/${TAG}_launches_bench_default.csv \
   python bench.py --steps 3 --warmup 3 --substeps 16 --e2e-substeps 16 --no-cpu-baseline --no-modes > gpurun_out/${TAG}_launches_bench.log 2>&1
# plain timings of the same commands (never read a number from a run under ncu)
for args in "16384 16 euler" "8192 16 euler1" "8192 8 rk4lap4" "512 512 rk4lap4" "512 2048 euler"; do
  timeout 200 python tools/rk_probe.py $args 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
done
for args in "16384 16 euler" "8192 8 rk4lap4"; do
  YH_ARITH=fast timeout 200 python tools/rk_probe.py $args 2>&1 | grep rk_probe | tee -a gpurun_out/${TAG}_probe.txt
done
