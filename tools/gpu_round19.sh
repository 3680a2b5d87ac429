#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_aux.py -m gpu -q -x -k "symmetry" 2>&1 | tail -5
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
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rd_rk_stream -s 4 -c 1 -o gpurun_out/r1_rd_rk_stream_8192_after_rotation -f \
   python bench.py --mode rk4lap4 --nx 8192 --ny 8192 --steps 1 --warmup 3 --substeps 2 --e2e-substeps 2 --no-cpu-baseline > gpurun_out/r1_rk_prof.log 2>&1
tail -2 gpurun_out/r1_rk_prof.log
