#!/usr/bin/env bash
set -x
mkdir -p gpurun_out/golden
python -m tests.golden.make_golden contour 2>&1 | tail -5
cp tests/golden/contour_*.npz gpurun_out/golden/
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 600 python tools/config_compare.py 2>&1 | grep '^{' > gpurun_out/config_compare.jsonl
python - <<PY
import json
for l in open("gpurun_out/config_compare.jsonl"):
    d=json.loads(l); print(d['config'][:60], '| ours', round(d['ours_Gcell_s'],2), '| ref', round(d.get('ref_Gcell_s',0),2), '| x', round(d.get('speedup', d.get('speedup_vs_sequential_reference',0)),2), '| bitwise', d.get('bitwise_vs_reference_nofma'))
PY
# exact vs FMA-contracted build of the same kernels (experiment)
for lib in lib lib_fma; do
  YH_LIB_PATH=$PWD/yolohtli_b200/$lib/libyolohtli_b200.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-substeps 64 2>/dev/null | tail -1 > gpurun_out/bench_euler_$lib.json
  YH_LIB_PATH=$PWD/yolohtli_b200/$lib/libyolohtli_b200.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline --mode rk4lap4 --nx 8192 --ny 8192 --substeps 16 --e2e-substeps 16 2>/dev/null | tail -1 > gpurun_out/bench_rk4_$lib.json
  python -c "
import json
for f in ('euler','rk4'):
    d=json.load(open('gpurun_out/bench_%s_$lib.json'%f)); print('$lib',f,round(d['value'],1),d['config']['workload'][:50])
"
done
