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
        if solid is None:   # the as-shipped loop: electrode probe after every step (main.cu:1040)
            sim = yh.Sim(p)
            sim.set_state(u0, v0)
            sim.run(256, trace=True)
            t0 = time.perf_counter()
            sim.run(nsteps, trace=True)
            rec["ours_traced_Gcell_s"] = nx * nx * nsteps / (time.perf_counter() - t0) / 1e9
            sim.close()
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
    # ---- C3: 512^2 symmetry reduction (BFECC + phase conditions + tips every step) -------------
    nx = 512
    pw = yh.default_params(nx, nx, timeIntOrder=1, lap4=0)
    sim = yh.Sim(pw)
    sim.cross_field_ic()
    sim.run(12001, tb_steps=4)          # warm-up: let the cross-field IC curl into a spiral
    tips = sim.tips()
    u0, v0 = sim.get_state()
    sim.close()
    tx, ty = (float(tips[-1]["x"]), float(tips[-1]["y"])) if len(tips) else (nx / 2.0, nx / 2.0)
    p = yh.default_params(nx, nx, reduce_sym=True, tipx0=tx, tipy0=ty)
    nsr = 1000
    sim = yh.Sim(p)
    sim.set_state(u0, v0)
    sim.run_sr(20, record=False)
    sim.set_state(u0, v0)
    c0_, phi0_ = np.zeros(3), np.zeros(3)
    import ctypes as C
    yh.lib().yh_sim_sr_state(sim._h, (C.c_double * 3)(*c0_), (C.c_double * 3)(*phi0_), 1)
    t0 = time.perf_counter()
    rec = sim.run_sr(nsr)
    t_host = (time.perf_counter() - t0) * 1e3
    # the same steps with the drift solve resident on the device (no host sync inside a step)
    sim.set_state(u0, v0)
    yh.lib().yh_sim_sr_state(sim._h, (C.c_double * 3)(*c0_), (C.c_double * 3)(*phi0_), 1)
    sim.run_sr(1, record=False)
    t0 = time.perf_counter()
    rec_dev = sim.run_sr_device(nsr - 1)
    t_dev = (time.perf_counter() - t0) * 1e3 * nsr / (nsr - 1)
    sim.close()
    t_ours = min(t_host, t_dev)
    r3 = {"config": "C3 512^2 symmetry reduction, default RK4+lap4 + tips + integrals + BFECC per step",
          "steps": nsr, "tips_at_start": int(len(tips)), "tip0": [tx, ty],
          "ours_Gcell_s": nx * nx * nsr / t_ours / 1e6, "ours_ms": t_ours, "ours_us_per_step": t_ours * 1e3 / nsr,
          "host_solve_us_per_step": t_host * 1e3 / nsr, "device_solve_us_per_step": t_dev * 1e3 / nsr,
          "device_vs_host_c_rel_diff": float(np.abs(rec_dev[:, :3] - rec[1:, :3]).max() / np.abs(rec[:, :3]).max()),
          "note": "wall clock; host_solve = one host sync per step for the 3x3 solve, device_solve = none"}
    if oracle_lib.have_reference():
        ref = oracle_lib.Reference(nofma=False)
        ref.init(p)
        ref.sr_run(u0[0], v0[0], 20)
        ru, rv, rrec, rms = ref.sr_run(u0[0], v0[0], nsr)
        r3.update(ref_Gcell_s=nx * nx * nsr / rms / 1e6, ref_ms=rms, speedup=rms / t_ours,
                  c_trace_rms_rel_diff=float(np.sqrt(((rec[:, :3] - rrec[:, :3]) ** 2).mean()) /
                                             max(1e-30, np.sqrt((rrec[:, :3] ** 2).mean()))),
                  c_final_ours=rec[-1, :3].tolist(), c_final_ref=rrec[-1, :3].tolist())
    print(json.dumps(r3), flush=True)
    out.append(r3)

    # ---- C5: batched restitution sweep, 32 sheets of 512^2 on this GPU (256 over 8 GPUs) --------
    nsim, nst = 32, 600
    p = yh.default_params(nx, nx, timeIntOrder=1, lap4=0)
    periods = np.linspace(600.0, 100.0, 256)[:nsim] / p.dt
    periods = periods.astype(np.int32)
    dur = int(10.0 / p.dt)
    area = synth.stim_area_square(nx, nx)
    sim = yh.Sim(p, n_sims=nsim)
    sim.set_pacing(periods, dur)
    sim.run_apd(20, stim_area=area)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sim.run_apd(nst, stim_area=area)
    t_ours = (time.perf_counter() - t0) * 1e3
    sim.close()
    r5 = {"config": "C5 batched sweep: 32 x 512^2 sheets per GPU, Euler+5pt, paced, sAPD every step",
          "steps": nst, "sheets": nsim, "ours_Gcell_s": nsim * nx * nx * nst / t_ours / 1e6, "ours_ms": t_ours}
    if oracle_lib.have_reference():
        ref = oracle_lib.Reference(nofma=False)
        ref.init(p)
        z = np.zeros((nx, nx))
        ref.apd_run(z, z, 20, int(periods[0]), dur, area)
        _, _, _, _, rms = ref.apd_run(z, z, nst, int(periods[0]), dur, area)
        _, _, _, _, rms_ship = ref.apd_run(z, z, nst, int(periods[0]), dur, area, mode=1)
        r5.update(ref_one_sheet_ms=rms, ref_Gcell_s=nx * nx * nst / rms / 1e6,
                  ref_as_shipped_Gcell_s=nx * nx * nst / rms_ship / 1e6,
                  speedup_vs_sequential_reference=nsim * rms / t_ours)
    print(json.dumps(r5), flush=True)
    out.append(r5)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "config_compare.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
