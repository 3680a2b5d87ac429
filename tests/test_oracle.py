"""CPU: pin the plain-C oracle against hand-derived cases and domain properties
(the reference ships no tests or golden vectors), and against the
golden fixtures produced by the reference's own kernels (tests/golden/)."""
import itertools
import os

import numpy as np
import pytest

from yolohtli_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rand_fields(nx, ny, seed=0):
    rng = np.random.default_rng(seed)
    return rng.uniform(-0.1, 1.1, (ny, nx)), rng.uniform(0.0, 1.0, (ny, nx))


def no_reaction(p):
    p.mu = 0.0
    p.eps = 0.0
    return p


ALL_MODES = []
for order, lap4, neu, gd in itertools.product((1, 2, 4), (0, 4), (1, 0), (1, 0)):
    ALL_MODES.append(dict(timeIntOrder=order, lap4=lap4, neumannBC=neu, gateDiff=gd))


def test_constant_field_is_fixed_point_of_diffusion(oracle):
    for m in ALL_MODES:
        if not m["neumannBC"]:
            continue
        p = no_reaction(oracle.params_default(40, 24, **m))
        u = np.full((24, 40), 0.37)
        v = np.zeros((24, 40))
        uo, vo = oracle.rd_step(p, u, v)
        assert np.abs(uo - 0.37).max() < 1e-15 and np.abs(vo).max() == 0.0


def test_neumann_diffusion_conserves_mass_in_the_interior_sense(oracle):
    # mirror BC: sum with boundary cells weighted 1/2 (trapezoid) is conserved by the 5-point step
    p = no_reaction(oracle.params_default(32, 20, timeIntOrder=1, lap4=0))
    u, v = rand_fields(32, 20, 1)
    v[:] = 0.0
    w = np.ones((20, 32)); w[0] *= .5; w[-1] *= .5; w[:, 0] *= .5; w[:, -1] *= .5
    uo, _ = oracle.rd_step(p, u, v)
    assert abs((uo * w).sum() - (u * w).sum()) < 1e-11


def test_euler_5pt_interior_cell_by_hand(oracle):
    p = oracle.params_default(8, 8, timeIntOrder=1, lap4=0)
    u, v = rand_fields(8, 8, 2)
    uo, vo = oracle.rd_step(p, u, v)
    i, j = 3, 4
    uu, vv = u[j, i] + 0.0, v[j, i] + 0.0
    I_sum = -(p.mu * uu * (1.0 - uu) * (uu - p.alpha) - uu * vv) - 0.0
    I_v = -(p.eps * (p.delta * (uu - p.gamma) * (p.beta - uu) - vv - p.theta))
    du = ((u[j, i - 1] - 2.0 * uu + u[j, i + 1]) * p.rx + (u[j + 1, i] - 2.0 * uu + u[j - 1, i]) * p.ry)
    dv = ((v[j, i - 1] - 2.0 * vv + v[j, i + 1]) * p.rx * p.rscale
          + (v[j + 1, i] - 2.0 * vv + v[j - 1, i]) * p.ry * p.rscale)
    du -= p.dt * I_sum
    dv -= p.dt * I_v
    assert uo[j, i] == u[j, i] + p.tc * (0.0 + 1.0 * du)
    assert vo[j, i] == v[j, i] + p.tc * (0.0 + 1.0 * dv)
    # mirror at the corner (helper_functions.cu:69-79): W of i=0 is i=1, S of j=0 is j=1
    uu = u[0, 0] + 0.0
    du0 = ((u[0, 1] - 2.0 * uu + u[0, 1]) * p.rx + (u[1, 0] - 2.0 * uu + u[1, 0]) * p.ry)
    vv = v[0, 0] + 0.0
    I0 = -(p.mu * uu * (1.0 - uu) * (uu - p.alpha) - uu * vv) - 0.0
    assert uo[0, 0] == u[0, 0] + p.tc * (0.0 + 1.0 * (du0 - p.dt * I0))


def test_fast_sweep_equals_generic_step(oracle):
    p = oracle.params_default(50, 34, timeIntOrder=1, lap4=0)
    u, v = rand_fields(50, 34, 3)
    a = oracle.rd_advance(p, 3, u, v, stim_mouse=True, point=(20, 15))   # fused sweep
    b_u, b_v = u, v
    for _ in range(3):
        b_u, b_v = oracle.rd_step(p, b_u, b_v, stim_mouse=True, point=(20, 15))
    assert np.array_equal(a[0], b_u) and np.array_equal(a[1], b_v)


def test_rk_stage_weights_and_velTan(oracle):
    # linear problem (no reaction, no gate diffusion of v): RK4 of u' = L u is the degree-4
    # Taylor polynomial applied step by step, with the reference's truncated literal weights
    p = no_reaction(oracle.params_default(16, 12, timeIntOrder=4, lap4=0, gateDiff=1))
    u, v = rand_fields(16, 12, 4)
    v[:] = 0.0   # with mu = eps = 0 and v = 0 the step is pure diffusion of u
    uo, vo, vtu, vtv = oracle.rd_step(p, u, v, velTan=True)
    p1 = no_reaction(oracle.params_default(16, 12, timeIntOrder=1, lap4=0, gateDiff=1))
    L = lambda x: oracle.rd_step(p1, x, np.zeros_like(x))[0] - x   # dt*L x
    k1 = L(u); k2 = L(u + .5 * k1); k3 = L(u + .5 * k2); k4 = L(u + k3)
    want = u + (0.166666666666667 * k1 + 0.333333333333333 * k2 + 0.333333333333333 * k3
                + 0.166666666666667 * k4)
    assert np.abs(uo - want).max() < 1e-14
    assert np.allclose(vtu.reshape(u.shape), (uo - u) / p.dt, atol=1e-10)


def test_solid_coefficient_table_all_32_patterns(oracle):
    # reactionDiffusion.cu:162-169 for every (sc, sw, se, sn, ss)
    p = no_reaction(oracle.params_default(8, 8, timeIntOrder=1, lap4=0, solidSwitch=1))
    u, v = rand_fields(8, 8, 5)
    i, j = 3, 3
    for sc, sw, se, sn, ss in itertools.product((0, 1), repeat=5):
        solid = np.ones((8, 8), dtype=np.uint8)
        solid[j, i], solid[j, i - 1], solid[j, i + 1], solid[j + 1, i], solid[j - 1, i] = sc, sw, se, sn, ss
        uo, _ = oracle.rd_step(p, u, v, solid=solid)
        cW = 1.0 if (sw and se and sc) else (2.0 if (sw and sc) else 0.0)
        cC = (2.0 if (sw or se) else 0.0) if sc else 0.0
        cE = 1.0 if (sw and se and sc) else (2.0 if (sc and se) else 0.0)
        cN = 1.0 if (sn and ss and sc) else (2.0 if (sn and sc) else 0.0)
        cCy = (2.0 if (sn or ss) else 0.0) if sc else 0.0
        cS = 1.0 if (sn and ss and sc) else (2.0 if (sc and ss) else 0.0)
        uu = u[j, i] + 0.0
        du = ((cW * u[j, i - 1] - cC * uu + cE * u[j, i + 1]) * p.rx
              + (cN * u[j + 1, i] - cCy * uu + cS * u[j - 1, i]) * p.ry)
        du -= p.dt * (-(0.0 * uu * (1.0 - uu) * (uu - p.alpha) - uu * (v[j, i] + 0.0)) - 0.0)
        want = (u[j, i] + p.tc * (0.0 + du)) if sc else 0.0
        assert uo[j, i] == want, (sc, sw, se, sn, ss)
        assert (uo[solid == 0] == 0.0).all()   # masked cells are exactly 0.0 (:521-535)


def test_mouse_stimulus_disc(oracle):
    p = oracle.params_default(64, 64, timeIntOrder=1, lap4=0)
    z = np.zeros((64, 64))
    uo, _ = oracle.rd_step(p, z, z, stim_mouse=True, point=(30, 28))
    jj, ii = np.nonzero(uo)
    # du = -dt*I_sum = dt*24.7 inside r^2 < 400 (reactionDiffusion.cu:61,134); diffusion of 0 is 0
    assert ((ii - 30) ** 2 + (jj - 28) ** 2 < 400).all()
    assert len(ii) == sum(1 for a in range(64) for b in range(64) if (a - 30) ** 2 + (b - 28) ** 2 < 400)
    assert np.allclose(uo[jj, ii], p.dt * 24.7)


def test_tip_known_answer(oracle):
    # u_present = Uth on the line x = 10.25, u_past = Uth on the line y = 7.5, both exactly
    # bilinear with a genuine xy term (the closed form divides by it): unique tip (10.25, 7.5).
    nx = ny = 24
    p = oracle.params_default(nx, ny)
    X, Y = np.meshgrid(np.arange(nx, dtype=float), np.arange(ny, dtype=float))
    present = p.Uth + 0.1 * (X - 10.25) + 0.01 * (X - 10.25) * (Y - 7.5)
    past = p.Uth + 0.1 * (Y - 7.5) + 0.02 * (X - 10.25) * (Y - 7.5)
    tips = oracle.tip_track(p, past, present, t=1.5)
    assert len(tips) == 1
    assert abs(tips[0]["x"] - 10.25) < 1e-5 and abs(tips[0]["y"] - 7.5) < 1e-5 and tips[0]["t"] == 1.5
    # Newton variant agrees (tipTracker.cu:244-432) when the window covers the cell
    tips2 = oracle.tip_track(p, past, present, algorithm=2)
    assert len(tips2) == 1 and abs(tips2[0]["x"] - 10.25) < 1e-5
    # canonical order: ascending cell index
    present2 = p.Uth + 0.1 * np.sin((X - 3.3) * 0.9) + 0.03 * np.cos(0.5 * Y + 0.4 * X)
    t3 = oracle.tip_track(p, past, present2)
    keys = np.floor(t3["x"]) + nx * np.floor(t3["y"])
    assert len(t3) >= 2 and (np.diff(keys) >= 0).all()


def test_bfecc_constant_and_translation(oracle):
    p = oracle.params_default(48, 40, reduce_sym=True)
    ax, ay = oracle.cxy_field(p, [0.3, -0.2, 0.0], [0.0, 0.0, 0.4])
    c = np.full((40, 48), 0.8)
    uo, vo = oracle.advect_bfecc(p, c, c * 0.5, ax, ay)
    assert np.abs(uo - 0.8).max() < 1e-15 and np.abs(vo - 0.4).max() < 1e-15
    # BFECC is exact to second order for a linear ramp advected by a uniform field
    X, Y = np.meshgrid(np.arange(48, dtype=float), np.arange(40, dtype=float))
    ramp = 0.01 * X + 0.02 * Y
    uo, _ = oracle.advect_bfecc(p, ramp, ramp, ax, ay)
    cx, cy = -ax.reshape(40, 48)[20, 20], -ay.reshape(40, 48)[20, 20]
    want = ramp - p.dt * (cx * 0.01 / p.hx + cy * 0.02 / p.hy)
    assert np.abs(uo - want)[5:-5, 5:-5].max() < 1e-12


def test_integrals_fused_equals_slice_then_trapz(oracle):
    p = oracle.params_default(96, 96, reduce_sym=True, tipOffsetX=30, tipOffsetY=30, tipx0=50.0, tipy0=44.0)
    u, v = rand_fields(96, 96, 7)
    vtu, vtv = rand_fields(96, 96, 8)
    ax, ay = oracle.cxy_field(p, [0.1, 0.2, 0.05], [0, 0, 0.3])
    s, s0 = oracle.slice(p, u, v, ax, ay)
    a = oracle.trapz(p, s, s0, vtu, vtv)
    b = oracle.sr_integrals(p, u, v, vtu, vtv, ax, ay)
    assert np.array_equal(a, b)
    # outside the disc the tangent fields are exactly zero (symmetryReduction.cu:171-182)
    X, Y = np.meshgrid(np.arange(96), np.arange(96))
    out = ((X - 50) ** 2 + (Y - 44) ** 2) >= 900
    assert all((np.asarray(q).reshape(96, 96)[out] == 0).all() for q in s + s0)
    # plain summation agrees to rounding (the reference's order is nondeterministic)
    ux0, vx0, ux, vx = s0[0], s0[3], s[0], s[3]
    want = p.hx * p.hy * (ux0 * ux + vx0 * vx).sum()
    assert abs(a[0] - want) <= 1e-12 * max(1.0, abs(want))
    # disc centre follows the last tip when count != 0
    tips = np.zeros(2, dtype=[("x", "<f4"), ("y", "<f4"), ("vx", "<f4"), ("vy", "<f4"), ("t", "<f4")])
    tips[1]["x"], tips[1]["y"] = 40.4, 60.6
    c = oracle.sr_integrals(p, u, v, vtu, vtv, ax, ay, tips=tips, count=5)
    p2 = oracle.params_default(96, 96, reduce_sym=True, tipOffsetX=30, tipOffsetY=30, tipx0=40.0, tipy0=61.0)
    d = oracle.sr_integrals(p2, u, v, vtu, vtv, ax, ay)
    assert np.array_equal(c, d)


def test_sapd_state_machine(oracle):
    # one cell going up through 0.15 at step 3 and down at step 9, twice
    p = oracle.params_default(4, 4)
    trace = [0.0, 0.05, 0.1, 0.3, 0.9, 0.9, 0.8, 0.5, 0.2, 0.1, 0.0, 0.05, 0.1, 0.3, 0.9, 0.8, 0.1, 0.0]
    seq = [np.full((4, 4), t) for t in trace]
    st = oracle.sapd_sequence(p, seq, count0=1)
    f1 = p.dt * (3 - (0.3 - 0.15) / (0.3 - 0.1))       # up-crossing between frames 2 -> 3, count = 3
    b1 = p.dt * (9 - (0.1 - 0.15) / (0.1 - 0.2))
    assert abs(st["APD1"][0] - (b1 - f1)) < 1e-15
    assert st["APD2"][0] > 0 and st["first"][0] == 0
    assert st["sAPD"][0] in (1.0, -1.0)


def test_contour_known_answers(oracle):
    """countour_kernel (spaceAPD.cu:18-153) by hand on a 5 x 4 sheet."""
    nx, ny = 5, 4
    p = oracle.params_default(nx, ny)
    # mode 1 (space-APD sign field): one sign change between columns 1|2 on every row
    s = np.ones((ny, nx)); s[:, 2:] = -1.0
    pts, n, plot = oracle.contour(p, None, s, 1, t=2.5, plot=True)
    # v0 = +1, v1x = -1 -> ppx = i + 1/(1+1) = 1.5; ppy = j (v0 == v1y); one hit per row, row order.
    # i = nx-1 reads the next row's first cell (+1) but zpmx uses v0*v0 there: no hit.
    assert n == ny and [tuple(q) for q in pts] == [(1.5, float(j), 2.5) for j in range(ny)]
    assert plot.reshape(ny, nx)[:, 1].all() and plot.sum() == ny
    # `zpmx<0 || zpmy<0 && sc`: the mask gates only the y-crossing (precedence as written, :62)
    s = np.ones((ny, nx)); s[2:, :] = -1.0; s[:, 3:] *= -1.0
    area = np.zeros((ny, nx), dtype=np.uint8)
    pts, n = oracle.contour(p, None, s, 1, stimArea=area)
    assert n == ny and all(q["x"] == 2.5 for q in pts)            # x-crossings survive the mask
    area[:] = 1
    pts, n = oracle.contour(p, None, s, 1, stimArea=area)
    assert n == ny + nx - 1    # + the y-crossing of row 1 in every column; (2,1) has both, one point
    both = [q for q in pts if q["x"] == 2.5 and q["y"] == 1.5]
    assert len(both) == 1
    # mode 2: u < th1 and v crosses th2 (0.85) -> the cell's integer coordinates
    u = np.full((ny, nx), 0.5); v = np.full((ny, nx), 0.9); v[:, 3:] = 0.8
    u[0, 2] = 0.95                                                 # not below th1 = 0.8: no point
    pts, n = oracle.contour(p, u, v, 2, t=1.0)
    assert [tuple(q) for q in pts] == [(2.0, float(j), 1.0) for j in range(1, ny)]
    # mode 3: both crossings of one cell, `-th2` first; |V| < th3 + 0.1 gates everything
    u = np.full((ny, nx), 0.5); v = np.full((ny, nx), 0.85 + 0.85 + 0.05); v[:, 3:] = 0.85 + 0.85 - 0.05
    pts, n = oracle.contour(p, u, v, 3)
    assert n == 0                                                  # |V0| = 0.9 > th3 + 0.1 = 0.8
    v = np.full((ny, nx), 0.85 - 0.7 + 0.02); v[:, 3:] = 0.85 - 0.7 - 0.02   # V + th3 changes sign
    pts, n = oracle.contour(p, u, v, 3, t=0.5)
    assert [tuple(q) for q in pts] == [(2.0, float(j), 0.5) for j in range(ny)]
    # capacity clips the stored list, not the count
    pts, n = oracle.contour(p, u, v, 3, capacity=2)
    assert n == ny and len(pts) == 2
    # colouring: get_rgba_kernel index = (int)((float)frac*(float)ncol), clamped; lines blank a pixel
    cm = np.arange(10, dtype=np.uint32) + 100
    f = np.array([[-0.1, 0.5, 1.0999, 1.1, 5.0], [-3.0, 0.02, 0.14, 0.26, 0.38]] * 2)
    ln = np.zeros((ny, nx), dtype=np.uint8); ln[0, 1] = 1
    out = oracle.rgba(p, f, cm, -0.1, 1.1, lines=ln).reshape(ny, nx)
    assert out[0].tolist() == [100, 0, 109, 109, 109] and out[1].tolist() == [100, 101, 102, 103, 104]


def test_hole_mask_generator_and_dat_roundtrip(tmp_path):
    m = synth.hole_mask(128, seed=7)
    frac = 1.0 - m.mean()
    assert 0.0184 < frac < 0.06 and m.dtype == np.uint8
    assert np.array_equal(m, synth.hole_mask(128, seed=7))
    f = tmp_path / "holes128.dat"
    synth.write_mask_dat(str(f), m)
    assert np.array_equal(synth.read_mask_dat(str(f), 128), m)   # main.cu:676-680 parse rule
    u, v = synth.fibrillation_ic(256, 256)
    assert set(np.unique(u)) == {0.0, 1.0} and u.shape == (256, 256)


@pytest.mark.parametrize("name", sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz")) if os.path.isdir(GOLDEN) else [])
def test_oracle_matches_reference_golden(oracle, name):
    """Fixtures written by tests/golden/make_golden.py on a B200 from the reference's own
    kernels (oracle/_ref, --fmad=false build)."""
    from tests.golden import make_golden
    g = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    make_golden.check_oracle(oracle, g)
