#!/usr/bin/env bash
# One gpurun call: GPU test suite, smoke(), default bench line.   gpurun --timeout 900 -- 'bash tools/gpu_s2_suite.sh TAG'
tag=${1:-s2}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > gpurun_out/${tag}_pytest_gpu.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2 > gpurun_out/${tag}_smoke.log
timeout 300 python bench.py 2>&1 | grep "^{" | tail -1 > gpurun_out/${tag}_bench_n1.json
cat gpurun_out/${tag}_pytest_gpu.log gpurun_out/${tag}_smoke.log; head -c 300 gpurun_out/${tag}_bench_n1.json
