#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rd_tile_rk|tip_kernel" -s 6 -c 2 -o gpurun_out/r1_small_rk_tip -f python tools/sr_profile.py 6 > gpurun_out/small_rk_tip.log 2>&1
tail -3 gpurun_out/small_rk_tip.log
