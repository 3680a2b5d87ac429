#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 900 python tools/config_compare.py 2>&1 | grep '^{' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['config'][:60], '| ours', round(d['ours_Gcell_s'],2), '| ref', round(d.get('ref_Gcell_s',0),2), '| x', round(d.get('speedup', d.get('speedup_vs_sequential_reference',0)),2), '| bitwise', d.get('bitwise_vs_reference_nofma'))
"
