"""Instruction mix of a kernel from an ncu report's source page (per-SASS-instruction executed
counts): opcode -> warp-level instructions executed, FP64 vs everything else, and the issue-slot
model of DESIGN.md (an FP64 instruction holds the dispatch port two cycles, tools/issue_mix.cu).

  python tools/ncu_mix.py gpurun_out/x.ncu-rep | x_source.csv [top=25]
"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
if rep.endswith(".csv"):      # an exported source page (ncu -i x.ncu-rep --page source --csv --print-source sass)
    txt = open(rep).read()
else:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
name = rows[0][1]
hdr = rows[1]
iS, iN, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, samples = Counter(), Counter()
for r in rows[2:]:
    if len(r) <= iN:
        continue
    toks = r[iS].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.rstrip(";").split(".")[0]
    ops[op] += int(r[iN] or 0)
    samples[op] += int(r[iSm] or 0)
tot = sum(ops.values())
fp64 = sum(v for k, v in ops.items() if k in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX"))
print(name)
print(f"warp instructions executed: {tot:,}   FP64: {fp64:,} ({100 * fp64 / tot:.1f} %)   other: {tot - fp64:,}")
print(f"issue-slot model: 2*FP64 + other = {2 * fp64 + tot - fp64:,}")
for k, v in ops.most_common(top):
    print(f"  {k:10s} {v:14,d}  {100 * v / tot:5.1f} %   samples {samples[k]}")
