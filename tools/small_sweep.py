"""Sweep strip width / chunk height for small sheets (latency-bound regime)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolohtli_b200 as yh  # noqa: E402
from yolohtli_b200 import host, synth  # noqa: E402


def timeit(p, u0, v0, nsteps, tb):
    uA, vA = torch.as_tensor(u0).cuda(), torch.as_tensor(v0).cuda()
    uB, vB = torch.zeros_like(uA), torch.zeros_like(vA)
    host.rd_advance(p, 8, uA, vA, uB, vB, tb_steps=tb)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        host.rd_advance(p, nsteps, uA, vA, uB, vB, tb_steps=tb, flags=1)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3 / nsteps   # us per time step


for nx in (512, 1024, 2048):
    u0, v0 = synth.cross_field_ic(nx, nx)
    for mode, kw, tbs, wvar, ws in (("euler", dict(timeIntOrder=1, lap4=0), (2, 4), "YH_FAST_W", (64, 128, 256)),
                                    ("rk4lap4", dict(), (1,), "YH_RK_W", (64, 128, 192))):
        p = yh.default_params(nx, nx, scale_L=True, **kw)
        for tb in tbs:
            res = []
            for w in ws:
                for ry in (0, 4, 8, 16, 32, 64):
                    os.environ[wvar] = str(w)
                    ryvar = "YH_FAST_RY" if mode == "euler" else "YH_RK_RY"
                    if ry:
                        os.environ[ryvar] = str(ry)
                    else:
                        os.environ.pop(ryvar, None)
                    us = timeit(p, u0, v0, 400 if mode == "euler" else 100, tb)
                    res.append((us, w, ry))
            res.sort()
            print(f"nx={nx} {mode} tb={tb}: best " + ", ".join(f"W={w} RY={ry}: {us:.2f}us" for us, w, ry in res[:4])
                  + f" | Gcell/s best {nx*nx/res[0][0]/1e3:.1f}", flush=True)
            auto = [r for r in res if r[2] == 0]
            print("    auto RY: " + ", ".join(f"W={w}: {us:.2f}us" for us, w, ry in auto), flush=True)
