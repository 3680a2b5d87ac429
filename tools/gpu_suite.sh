#!/usr/bin/env bash
# One gpurun call on a single B200: the GPU test suite, smoke(), the default bench line and the
# reference arm.  Logs land in gpurun_out/ (merged back by gpurun).
#   gpurun --timeout 780 -- 'bash tools/gpu_suite.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
# new this session first, verbosely: a failure here must not hide the state of the rest
timeout 240 python -m pytest tests/test_gpu_sr_slabs.py -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_sr_slabs.log
timeout 420 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_sr_slabs.py 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 300 python bench.py 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_default.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print("bench", round(d["value"], 1), d["unit"], "| e2e", round(d["e2e"]["value"], 1), "| roofline frac", round(d["roofline"]["frac"], 3),
      "| cpu", round(d["cpu_baseline"]["value"], 3), "on", d["cpu_baseline"]["cores"], "cores | clocks", d["clocks"])
PY
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_default_ref.json
python -c "import json; d=json.load(open('gpurun_out/bench_default_ref.json')); print('reference arm', round(d['value'],2), d['unit'])"
