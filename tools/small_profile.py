"""Small-sheet regime: wall time per launch (CUDA events over a long loop) for the kernels that
serve C1 / C2; run once plain and once under `ncu --metrics gpu__time_duration.sum` to split
launch overhead from kernel time."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolohtli_b200 as yh  # noqa: E402
from yolohtli_b200 import host, synth  # noqa: E402

short = len(sys.argv) > 1 and sys.argv[1] == "short"


def timeit_graph(p, u0, v0, nsteps, tb, solid=None, chunk=64):
    """Same loop replayed from a CUDA graph of `chunk` time steps (even number of launches, so the
    ping-pong returns to the A buffers)."""
    uA, vA = torch.as_tensor(u0).cuda(), torch.as_tensor(v0).cuda()
    uB, vB = torch.zeros_like(uA), torch.zeros_like(vA)
    ds = torch.as_tensor(solid).cuda() if solid is not None else None
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        host.rd_advance(p, chunk, uA, vA, uB, vB, tb_steps=tb, solid=ds)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        ru, rv = host.rd_advance(p, chunk, uA, vA, uB, vB, tb_steps=tb, flags=1, solid=ds)
    assert ru is uA
    reps = max(1, nsteps // chunk)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3 / (reps * chunk)


def timeit(p, u0, v0, nsteps, tb, solid=None):
    uA, vA = torch.as_tensor(u0).cuda(), torch.as_tensor(v0).cuda()
    uB, vB = torch.zeros_like(uA), torch.zeros_like(vA)
    ds = torch.as_tensor(solid).cuda() if solid is not None else None
    host.rd_advance(p, 8, uA, vA, uB, vB, tb_steps=tb, solid=ds)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(1 if short else 3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        host.rd_advance(p, nsteps, uA, vA, uB, vB, tb_steps=tb, flags=1, solid=ds)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3 / nsteps   # us per time step


for nx in (512, 1024):
    u0, v0 = synth.cross_field_ic(nx, nx)
    mask = synth.hole_mask(nx, seed=1)
    for name, kw, tb, solid, n in (("euler tb4", dict(timeIntOrder=1, lap4=0), 4, None, 2000),
                                   ("euler tb1", dict(timeIntOrder=1, lap4=0), 1, None, 1000),
                                   ("rk4lap4", dict(), 1, None, 400),
                                   ("euler tb4 holes", dict(timeIntOrder=1, lap4=0, solidSwitch=1), 4, mask, 2000),
                                   ("rk4 holes", dict(solidSwitch=1), 1, mask, 400)):
        p = yh.default_params(nx, nx, scale_L=True, **kw)
        for path in ("stream", "tile"):
            os.environ["YH_RD_PATH"] = path
            us = timeit(p, u0, v0, 40 if short else n, tb, solid)
            print(f"nx={nx} {name:16s} {path:6s}: {us:7.2f} us/step  {us * tb:7.2f} us/launch  {nx * nx / us / 1e3:7.1f} Gcell/s", flush=True)
            if not short and solid is None:
                ug = timeit_graph(p, u0, v0, n, tb, solid)
                print(f"nx={nx} {name:16s} {path:6s}: {ug:7.2f} us/step  {ug * tb:7.2f} us/launch  {nx * nx / ug / 1e3:7.1f} Gcell/s  [CUDA graph, 64 steps]", flush=True)
