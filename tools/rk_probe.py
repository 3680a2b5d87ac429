"""Default-mode (RK4 + 4th-order Laplacian) step: device time per step and a bit-level checksum of the
result, for A/B runs of kernel variants (environment switches are read by the library at launch).

  python tools/rk_probe.py [n=8192] [steps=12] [mode=rk4lap4|rk4|rk2lap4|rk4holes|euler|euler1|euler2]
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolohtli_b200 as yh  # noqa: E402
from yolohtli_b200 import host, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
mode = sys.argv[3] if len(sys.argv) > 3 else "rk4lap4"
over = {"rk4lap4": {}, "rk4": dict(lap4=0), "rk2lap4": dict(timeIntOrder=2), "rk4holes": dict(solidSwitch=1),
        "euler": dict(timeIntOrder=1, lap4=0), "euler1": dict(timeIntOrder=1, lap4=0),
        "euler2": dict(timeIntOrder=1, lap4=0)}[mode]
tb = {"euler1": 1, "euler2": 2}.get(mode, 0)          # time steps per HBM pass (0 = the library's choice)
p = yh.default_params(n, n, scale_L=True, **over)
u0, v0 = synth.fibrillation_ic(n, n) if n >= 1024 else synth.cross_field_ic(n, n)
solid = None
if mode == "rk4holes":      # C2: blood-vessel obstacles (the mask branch has no 4th-order terms)
    m = synth.hole_mask(n)
    u0, v0 = u0 * m, v0 * m
    solid = torch.as_tensor(m).cuda()
uA, vA = torch.as_tensor(u0).cuda(), torch.as_tensor(v0).cuda()
uB, vB = torch.zeros_like(uA), torch.zeros_like(vA)
ru, rv = host.rd_advance(p, 4, uA, vA, uB, vB, solid=solid, tb_steps=tb)          # warm-up, 4 steps: result back in A
torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ru, rv = host.rd_advance(p, steps, ru, rv, uB if ru is uA else uA, vB if ru is uA else vA, flags=1, solid=solid, tb_steps=tb)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
h = hashlib.sha256(ru.cpu().numpy().tobytes() + rv.cpu().numpy().tobytes()).hexdigest()[:16]
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("YH_"))
print(f"rk_probe {mode} {n}x{n}: {best / steps * 1e3:9.1f} us/step  {n * n * steps / best / 1e6:7.2f} Gcell/s  "
      f"sha={h}  [{tag}]", flush=True)
