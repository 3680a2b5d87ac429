#!/usr/bin/env bash
# One gpurun call: GPU test suite, smoke(), default bench line, reference arm.
#   gpurun --timeout 900 -- 'bash tools/gpu_s2_suite.sh TAG'
tag=${1:-s2}
mkdir -p gpurun_out
(timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > gpurun_out/${tag}_pytest_gpu.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2 > gpurun_out/${tag}_smoke.log
timeout 300 python bench.py 2>&1 | grep "^{" | tail -1 > gpurun_out/${tag}_bench_n1.json
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | grep "^{" | tail -1 > gpurun_out/${tag}_bench_ref.json
cat gpurun_out/${tag}_pytest_gpu.log gpurun_out/${tag}_smoke.log; head -c 300 gpurun_out/${tag}_bench_n1.json
