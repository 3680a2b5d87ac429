"""Per-config numbers on one GPU: ours vs the reference's own kernels (oracle/_ref), with parity.
   C1 512^2 spiral (Euler+5pt and default RK4+lap4), C2 1024^2 holes + fibrillation IC."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolohtli_b200 as yh  # noqa: E402
from yolohtli_b200 import host, synth  # noqa: E402
from tests import oracle_lib  # noqa: E402


def ours(p, u0, v0, nsteps, solid=None, tb=0, reps=3):
    uA, vA = torch.as_tensor(u0).cuda(), torch.as_tensor(v0).cuda()
    uB, vB = torch.zeros_like(uA), torch.zeros_like(vA)
    ds = torch.as_tensor(solid).cuda() if solid is not None else None
    best = 1e9
    for r in range(reps):
        uA.copy_(torch.as_tensor(u0)); vA.copy_(torch.as_tensor(v0))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ru, rv = host.rd_advance(p, nsteps, uA, vA, uB, vB, tb_steps=tb, solid=ds)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return ru.cpu().numpy(), rv.cpu().numpy(), best


def main():
    out = []
    for name, nx, kw, ic, solid_seed, nsteps in [
        ("C1 512^2 spiral, Euler+5pt", 512, dict(timeIntOrder=1, lap4=0), "cross", None, 10000),
        ("C1 512^2 spiral, default RK4+lap4", 512, dict(), "cross", None, 2000),
        ("C2 1024^2 holes, fibrillation, Euler", 1024, dict(timeIntOrder=1, lap4=0, solidSwitch=1), "fib", 1, 2000),
        ("C2 1024^2 holes, fibrillation, default RK4", 1024, dict(solidSwitch=1), "fib", 1, 1000),
    ]:
        p = yh.default_params(nx, nx, scale_L=True, **kw)
        u0, v0 = synth.cross_field_ic(nx, nx) if ic == "cross" else synth.fibrillation_ic(nx, nx)
        solid = synth.hole_mask(nx, seed=solid_seed) if solid_seed is not None else None
        if solid is not None:
            u0, v0 = u0 * solid, v0 * solid
        gu, gv, ms = ours(p, u0, v0, nsteps, solid=solid)
        rec = {"config": name, "steps": nsteps, "ours_Gcell_s": nx * nx * nsteps / ms / 1e6, "ours_ms": ms}
        if oracle_lib.have_reference():
            ref = oracle_lib.Reference(nofma=False)
            ref.init(p)
            best = 1e9
            for _ in range(2):
                ru, rv, t = ref.rd_run(u0, v0, nsteps, solid=solid)
                best = min(best, t)
            _, _, t_ship = ref.rd_run(u0, v0, nsteps, solid=solid, mode=1)   # + per-step blocking D2H (as shipped)
            rec.update(ref_Gcell_s=nx * nx * nsteps / best / 1e6, ref_as_shipped_Gcell_s=nx * nx * nsteps / t_ship / 1e6,
                       speedup=best / ms, max_abs_diff_u=float(np.abs(gu - ru).max()))
            if p.timeIntOrder == 1:
                refn = oracle_lib.Reference(nofma=True)
                refn.init(p)
                nu, nv, _ = refn.rd_run(u0, v0, nsteps, solid=solid)
                rec["bitwise_vs_reference_nofma"] = bool(np.array_equal(gu, nu) and np.array_equal(gv, nv))
        print(json.dumps(rec), flush=True)
        out.append(rec)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "config_compare.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
