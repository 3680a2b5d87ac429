"""GPU: the BASELINE configurations at their stated sizes (C3 and C5; C1 lives in test_gpu_rd.py, C2 in
test_gpu_rd.py::test_masked_temporal_blocking..., C4 in test_full_size_16384_sheet).

C3  512 x 512 symmetry-reduction mode (main.cu:894-954): BFECC advection + trapz phase conditions + tip
    tracking of a developed, meandering spiral, integration disc of radius 160 cells, 2000 steps --
    bitwise against the plain-C composition of the same loop, and the drift / phase history and the
    final fields against the REFERENCE's own loop (its kernels race in this mode, so the bound is
    calibrated on five reference runs, SURVEY 8c tier T2).
C5  32 independent 512 x 512 paced sheets (one GPU's share of the 256-sheet sweep), sAPD every step --
    bitwise against the reference's own loop for every sheet."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from tests import oracle_lib  # noqa: E402
from yolohtli_b200 import synth  # noqa: E402

NX = 512


@pytest.fixture(scope="module")
def spiral(yh):
    """A developed spiral on the reference's default 512^2 sheet: 12001 Euler steps from the cross-field
    initial condition (the bitwise-pinned path), and the tip that centres the integration disc."""
    pw = yh.default_params(NX, NX, timeIntOrder=1, lap4=0)
    sim = yh.Sim(pw)
    sim.cross_field_ic()
    sim.run(12001, tb_steps=4)
    tips = sim.tips()
    u0, v0 = (a[0].copy() for a in sim.get_state())
    sim.close()
    assert len(tips) > 0, "the warm-up must leave a spiral tip"
    return u0, v0, float(tips[-1]["x"]), float(tips[-1]["y"])


def test_c3_symmetry_reduction_512_vs_oracle(oracle, yh, spiral):
    u0, v0, tx, ty = spiral
    nsteps = 2000
    p = yh.default_params(NX, NX, reduce_sym=True, tipx0=tx, tipy0=ty)
    assert p.tipOffsetX == 160 and p.tipOffsetY == 160
    sim = yh.Sim(p)
    sim.set_state(u0[None], v0[None])
    rec = sim.run_sr(nsteps)
    gu, gv = (a[0] for a in sim.get_state())
    gc, gphi = sim.sr_state()
    sim.close()
    # the same loop composed from oracle pieces (the composition of test_symmetry_reduction_step_loop)
    u, v = u0.copy(), v0.copy()
    c, phi = np.zeros(3), np.zeros(3)
    ax, ay = np.zeros(NX * NX), np.zeros(NX * NX)
    with_tips = 0
    for count in range(nsteps):
        us, vs, vtu, vtv = oracle.rd_step(p, u, v, velTan=True)
        tips = oracle.tip_track(p, us, u, t=p.dt * count)
        with_tips += len(tips) > 0
        assert np.array_equal(rec[count], np.concatenate([c, phi])), count
        I = oracle.sr_integrals(p, u, v, vtu, vtv, ax, ay, tips=tips, count=count)
        if count == 0:
            c = oracle.solve_matrix(c, phi, I)
            ax, ay = oracle.cxy_field(p, c, phi)
            I = oracle.sr_integrals(p, u, v, vtu, vtv, ax, ay, tips=tips, count=count)
        c = oracle.solve_matrix(c, phi, I)
        ax, ay = oracle.cxy_field(p, c, phi)
        u, v = oracle.advect_bfecc(p, us, vs, ax, ay)
        phi = np.array([phi[q] + c[q] * p.dt for q in range(3)])
    assert with_tips == nsteps, "the spiral tip must be tracked in every step"
    assert np.array_equal(gu, u) and np.array_equal(gv, v)
    assert np.array_equal(gc, c) and np.array_equal(gphi, phi)
    assert np.abs(rec[:, :3]).max() > 1e-3 and np.isfinite(rec).all()


@pytest.mark.skipif(not oracle_lib.have_reference(), reason="oracle/_ref not built")
def test_c3_symmetry_reduction_512_vs_reference_loop(yh, spiral):
    """(c, phi) history and final fields vs yref_sr_run = the reference's main.cu:894-954 loop with its
    own wrappers.  Its RK4 / lap4 / BFECC kernels race, so runs differ; bound = max(floor,
    3 x the spread of five reference runs), both printed."""
    u0, v0, tx, ty = spiral
    nsteps = 2000
    p = yh.default_params(NX, NX, reduce_sym=True, tipx0=tx, tipy0=ty)
    ref = oracle_lib.Reference(nofma=False)
    ref.init(p)
    runs = [ref.sr_run(u0, v0, nsteps) for _ in range(5)]
    rrec = np.stack([r[2] for r in runs])                       # [5, nsteps, 6]
    ru = np.stack([r[0] for r in runs])
    spread_rec = np.abs(rrec - rrec[0]).max(axis=0)             # [nsteps, 6]
    spread_u = float(np.abs(ru - ru[0]).max())
    for name, runner in (("host solve", "run_sr"), ("device-resident solve", "run_sr_device")):
        sim = yh.Sim(p)
        sim.set_state(u0[None], v0[None])
        rec = getattr(sim, runner)(nsteps)
        gu = sim.get_state()[0][0]
        sim.close()
        scale_c = np.abs(rrec[0][:, :3]).max()
        scale_phi = np.abs(rrec[0][:, 3:]).max()
        dc = np.abs(rec[:, :3] - rrec[0][:, :3]).max() / scale_c
        dphi = np.abs(rec[:, 3:] - rrec[0][:, 3:]).max() / scale_phi
        sc = spread_rec[:, :3].max() / scale_c
        sphi = spread_rec[:, 3:].max() / scale_phi
        du = float(np.abs(gu - ru[0]).max())
        print(f"C3 vs reference loop [{name}]: |c| diff {dc:.3e} (reference spread {sc:.3e}), |phi| diff {dphi:.3e} "
              f"(spread {sphi:.3e}), final u diff {du:.3e} (spread {spread_u:.3e}); c scale {scale_c:.3e}, phi scale {scale_phi:.3e}")
        # measured (B200, round 2): |c| 0.16 of scale (the reference's own runs spread by 0.17: c is a noisy
        # pointwise quantity), |phi| 4.0e-3 of scale (spread 1.6e-4), final u 2.6e-3 (spread 5.5e-4)
        assert dc <= max(0.05, 3 * sc) and dphi <= max(0.02, 3 * sphi)
        assert du <= max(0.01, 3 * spread_u)
    assert scale_c > 1e-3, "a meandering spiral must drift"


@pytest.mark.skipif(not oracle_lib.have_reference(), reason="oracle/_ref not built")
def test_c5_sweep_512x32_vs_reference_loop(yh):
    """One GPU's share of the 256-sheet restitution sweep at full size: 32 sheets of 512^2, stimulation
    periods as bench.py --workload sweep, sAPD every step; fields and APD maps of EVERY sheet bitwise
    equal to the reference's own loop (reference kernels built with --fmad=false: Euler + 5-point is
    race-free, so this tier is exact)."""
    nsim, nsteps = 32, 36000
    p = yh.default_params(NX, NX, timeIntOrder=1, lap4=0)
    periods = (np.linspace(600.0, 100.0, 256) / p.dt).astype(np.int32)[:nsim]
    dur = int(10.0 / p.dt)
    area = synth.stim_area_square(NX, NX)
    sim = yh.Sim(p, n_sims=nsim)
    z = np.zeros((nsim, NX, NX))
    sim.set_state(z, z)
    sim.set_pacing(periods, dur)
    sim.run_apd(nsteps, stim_area=area)
    gu, gv = sim.get_state()
    a1, a2 = sim.get_apd()
    sim.close()
    ref = oracle_lib.Reference(nofma=True)
    ref.init(p)
    z1 = np.zeros((NX, NX))
    done = 0
    for s in range(nsim):
        ru, rv, r1, r2, _ = ref.apd_run(z1, z1, nsteps, int(periods[s]), dur, area)
        assert np.array_equal(gu[s], ru) and np.array_equal(gv[s], rv), s
        assert np.array_equal(a1[s].ravel(), r1) and np.array_equal(a2[s].ravel(), r2), s
        done += int(np.abs(r1).max() > 0)
    assert done == nsim, "every sheet must have completed an action potential (APD1 set)"
    assert len({a1[s].tobytes() for s in range(nsim)}) > 1, "different pacing periods must give different sheets"
