#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
python tools/small_profile.py 2>&1 | grep "^nx\|Error\|error" | grep -v holes | tee gpurun_out/small_profile_graph.txt
