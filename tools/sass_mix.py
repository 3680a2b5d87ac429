"""Static opcode mix of one kernel in an object file: python tools/sass_mix.py <obj> <substring of mangled name>"""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
fn, mix = None, collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    if fn and pat in fn:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            mix[m.group(1)] += 1
tot = sum(mix.values())
fp = mix["DADD"] + mix["DMUL"] + mix["DFMA"]
print(f"{pat}: {tot} instructions, FP64 {fp} ({100 * fp / max(tot, 1):.0f}%)")
print("  " + "  ".join(f"{k}:{v}" for k, v in mix.most_common(18)))
