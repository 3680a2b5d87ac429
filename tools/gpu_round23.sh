#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rd.py -m gpu -q -x -s -k "default_mode_trace" 2>&1 | tail -12
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/sr_launches.csv python tools/sr_profile.py 12 > /dev/null 2>&1
python - <<'PY'
import csv, collections, re
rows = list(csv.reader(l for l in open("gpurun_out/sr_launches.csv") if l.startswith('"')))
h = rows[0]; ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
agg = collections.defaultdict(list)
for r in rows[-80:]:
    agg[(re.sub(r"\(.*", "", r[ki])[:70], r[gi])].append(float(r[vi].replace(",", "")) / 1e3)
for (k, g), v in sorted(agg.items()):
    print(f"{k:70s} grid {g:14s} n={len(v):4d} median {sorted(v)[len(v)//2]:8.2f} us")
PY
