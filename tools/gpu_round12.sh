#!/usr/bin/env bash
# GPU suite + per-config comparison with the small-sheet kernels forced on / off
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for path in stream tile; do
YH_RD_PATH=$path timeout 600 python tools/config_compare.py 2>&1 | grep '^{' > gpurun_out/config_compare_$path.jsonl
python - <<PY
import json
for l in open("gpurun_out/config_compare_$path.jsonl"):
    d=json.loads(l); print("$path", d['config'][:60], '| ours', round(d['ours_Gcell_s'],2), '| ref', round(d.get('ref_Gcell_s',0),2), '| x', round(d.get('speedup', d.get('speedup_vs_sequential_reference',0)),2), '| bitwise', d.get('bitwise_vs_reference_nofma'))
PY
done
