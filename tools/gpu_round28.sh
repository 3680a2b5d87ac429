#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/sr_launches.csv python tools/sr_profile.py 12 > /dev/null 2>&1
python - <<'PY'
import csv, collections, re
rows = list(csv.reader(l for l in open("gpurun_out/sr_launches.csv") if l.startswith('"')))
h = rows[0]; ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
agg = collections.defaultdict(list)
for r in rows[-60:]:
    agg[(re.sub(r"\(.*", "", r[ki])[:70], r[gi])].append(float(r[vi].replace(",", "")) / 1e3)
for (k, g), v in sorted(agg.items()):
    print(f"{k:70s} grid {g:14s} n={len(v):4d} median {sorted(v)[len(v)//2]:8.2f} us")
PY
timeout 600 python tools/config_compare.py 2>&1 | grep '^{' > gpurun_out/config_compare.jsonl
python - <<PY
import json
for l in open("gpurun_out/config_compare.jsonl"):
    d=json.loads(l); print(d['config'][:60], '| ours', round(d['ours_Gcell_s'],2), '| ref', round(d.get('ref_Gcell_s',0),2), '| x', round(d.get('speedup', d.get('speedup_vs_sequential_reference',0)),2), '| bitwise', d.get('bitwise_vs_reference_nofma'), {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items() if 'solve' in k or 'rel_diff' in k})
PY
