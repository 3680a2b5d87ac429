#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 600 python tools/config_compare.py 2>&1 | grep '^{' > gpurun_out/config_compare.jsonl
python - <<PY
import json
for l in open("gpurun_out/config_compare.jsonl"):
    d=json.loads(l); print(d['config'][:60], '| ours', round(d['ours_Gcell_s'],2), '| ref', round(d.get('ref_Gcell_s',0),2), '| x', round(d.get('speedup', d.get('speedup_vs_sequential_reference',0)),2), '| traced', round(d.get('ours_traced_Gcell_s',0),2), 'vs as-shipped', round(d.get('ref_as_shipped_Gcell_s',0),2))
PY
timeout 600 python bench.py --workload sweep --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_sweep_n1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('sweep', d['value'], d['e2e']['value'])"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
