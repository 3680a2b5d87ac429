"""GPU: the four-columns-per-thread RK4 kernel (csrc/rd_rkq.cu) and the FAST arithmetic flavour.

EXACT flavour: bit-identical to the plain-C oracle (and so to rd_rk.cu) in every configuration the kernel
serves -- with / without the 4th-order Laplacian, default and non-default model constants, velTan output,
row ranges of a slab, sheets whose first / last rows exercise the no-flux mirrors.

FAST flavour (YH_ARITH_FAST: stencil coefficients combined on the host, FMA chains -- the reference's
shipped build contracts FMAs as well, Makefile:9): the same update up to rounding.  Stated tolerances:
  * <= 1e-14 (absolute, fields are O(1)) after 25 default-mode steps (measured 4.4e-16; the reference's own
    FMA-contracted build is 1e-12 from the exact one), <= 1e-13 after 100 Euler steps (2.4e-15);
  * after 2000 steps of a rotating spiral the electrode voltages and the whole field stay within 1e-12 of
    the exact run (measured 1e-15 / 4.6e-15) and the spiral tip within 1e-4 cell (measured: identical) --
    ten orders below the reference's run-to-run spread in this mode (5e-5 in voltage, 0.01-0.15 cell in
    tip position: its stage kernels race)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from yolohtli_b200 import host, synth  # noqa: E402


@pytest.fixture(autouse=True)
def quad_kernel():
    os.environ["YH_RK_KERNEL"] = "quad"
    os.environ["YH_EULER_KERNEL"] = "quad"
    os.environ["YH_RD_PATH"] = "stream"
    yield
    for k in ("YH_RK_KERNEL", "YH_EULER_KERNEL", "YH_RD_PATH"):
        os.environ.pop(k, None)


@pytest.fixture
def fast(yh):
    assert yh.lib().yh_set_arithmetic(1) == 0
    yield
    assert yh.lib().yh_set_arithmetic(0) == 0


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def advance(p, n, u, v, rows=None):
    uA, vA = dev(u), dev(v)
    uB, vB = torch.zeros_like(uA), torch.zeros_like(vA)
    ru, rv = host.rd_advance(p, n, uA, vA, uB, vB, rows=rows)
    torch.cuda.synchronize()
    return ru.cpu().numpy(), rv.cpu().numpy()


def fields(nx, ny, seed=0):
    rng = np.random.default_rng(seed)
    return rng.uniform(-0.1, 1.1, (ny, nx)), rng.uniform(0.0, 1.0, (ny, nx))


@pytest.mark.parametrize("nx,ny", [(128, 40), (256, 96), (500, 131), (1024, 70), (2048, 33)])
@pytest.mark.parametrize("kw", [dict(), dict(lap4=0), dict(mu=1.1, delta=0.9, gamma=0.05, theta=0.01, tc=0.9),
                                dict(lap4=0, alpha=0.15, eps=0.01)])
def test_rk_quad_exact_bitwise_vs_oracle(oracle, nx, ny, kw):
    if nx % 4:
        nx -= nx % 4
    p = oracle.params_default(nx, ny, **kw)
    u, v = fields(nx, ny, seed=nx + ny)
    u[3:9, 5:40] = -0.0          # raw user data may hold -0.0
    want = oracle.rd_advance(p, 3, u, v)
    got = advance(p, 3, u, v)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert np.array_equal(np.signbit(got[0]), np.signbit(want[0]))


def test_rk_quad_veltan_and_row_ranges(oracle):
    """velTan = rhs / dt as the symmetry-reduction loop needs it, and a slab: rows [r0, r1) of a local array
    whose row 0 is global row jg0 (mirrors only at the GLOBAL edges)."""
    nx, ny = 384, 120
    p = oracle.params_default(nx, ny)
    u, v = fields(nx, ny, seed=5)
    wu, wv, wtu, wtv = oracle.rd_step(p, u, v, velTan=True)
    du, dv = dev(u), dev(v)
    uo, vo, tu, tv = (torch.zeros_like(du) for _ in range(4))
    host.rd_step(p, du, dv, uo, vo, velTan=(tu, tv))
    torch.cuda.synchronize()
    assert np.array_equal(uo.cpu().numpy(), wu) and np.array_equal(vo.cpu().numpy(), wv)
    assert np.array_equal(tu.cpu().numpy(), wtu) and np.array_equal(tv.cpu().numpy(), wtv)
    # slab: global rows [30, 100) stored with 4 ghost rows per side, one RK4 step on the owned rows
    lo, hi, H = 30, 100, 4
    q = p.copy()
    q.ny, q.ny_global, q.jg0 = hi - lo + 2 * H, ny, lo - H
    su, sv = u[lo - H:hi + H], v[lo - H:hi + H]
    gu, gv = advance(q, 1, su, sv, rows=(H, H + hi - lo))
    assert np.array_equal(gu[H:H + hi - lo], wu[lo:hi]) and np.array_equal(gv[H:H + hi - lo], wv[lo:hi])
    # ... and one that touches the top of the sheet (no ghost rows there, no-flux mirror instead)
    q.ny, q.jg0 = 60 + H, 0
    gu, gv = advance(q, 1, u[:60 + H], v[:60 + H], rows=(0, 60))
    assert np.array_equal(gu[:60], wu[:60]) and np.array_equal(gv[:60], wv[:60])


@pytest.mark.parametrize("kw", [dict(), dict(lap4=0), dict(mu=1.1, delta=0.9, gamma=0.05, theta=0.01)])
def test_fast_flavour_25_steps_within_1e12(oracle, yh, fast, kw):
    nx = ny = 512
    p = oracle.params_default(nx, ny, **kw)
    u, v = synth.cross_field_ic(nx, ny)
    u = u + 0.05 * np.sin(0.07 * np.arange(nx))[None, :] * np.cos(0.05 * np.arange(ny))[:, None]
    want = oracle.rd_advance(p, 25, u, v)
    got = advance(p, 25, u, v)
    eu, ev = np.abs(got[0] - want[0]).max(), np.abs(got[1] - want[1]).max()
    print(f"fast vs exact after 25 steps {kw}: max |du| {eu:.3e}, max |dv| {ev:.3e}")
    assert 0 < eu <= 1e-14 and ev <= 1e-14          # different rounding (so not 0), nothing more; measured 4.4e-16


@pytest.mark.parametrize("tb", [4, 2, 1])
@pytest.mark.parametrize("kw", [dict(), dict(gateDiff=0), dict(mu=1.1, delta=0.9, gamma=0.05, theta=0.01, tc=0.9)])
def test_fast_flavour_euler_100_steps_within_1e12(oracle, yh, fast, kw, tb):
    """The temporally blocked Euler kernel (rd_quad.cu) in the FAST flavour: 100 steps vs the exact oracle."""
    nx, ny = 512, 300
    p = oracle.params_default(nx, ny, timeIntOrder=1, lap4=0, **kw)
    u, v = synth.cross_field_ic(nx, ny)
    u = u + 0.05 * np.sin(0.07 * np.arange(nx))[None, :] * np.cos(0.05 * np.arange(ny))[:, None]
    want = oracle.rd_advance(p, 100, u, v)
    uA, vA = dev(u), dev(v)
    uB, vB = torch.zeros_like(uA), torch.zeros_like(vA)
    ru, rv = host.rd_advance(p, 100, uA, vA, uB, vB, tb_steps=tb)
    torch.cuda.synchronize()
    eu, ev = np.abs(ru.cpu().numpy() - want[0]).max(), np.abs(rv.cpu().numpy() - want[1]).max()
    print(f"Euler fast vs exact after 100 steps {kw} T={tb}: max |du| {eu:.3e}, max |dv| {ev:.3e}")
    assert 0 < eu <= 1e-13 and ev <= 1e-13          # measured 2.4e-15 / 5.3e-15


def test_fast_flavour_tile_kernel_25_steps(oracle, yh, fast):
    """The small-sheet (tile) RK4 + lap4 kernel in the FAST flavour: the reference's default 512^2 sheet."""
    os.environ["YH_RD_PATH"] = "tile"
    nx = ny = 512
    for kw in (dict(), dict(mu=1.1, delta=0.9, gamma=0.05, theta=0.01)):
        p = oracle.params_default(nx, ny, **kw)
        u, v = synth.cross_field_ic(nx, ny)
        u = u + 0.05 * np.sin(0.07 * np.arange(nx))[None, :] * np.cos(0.05 * np.arange(ny))[:, None]
        want = oracle.rd_advance(p, 25, u, v)
        got = advance(p, 25, u, v)
        eu, ev = np.abs(got[0] - want[0]).max(), np.abs(got[1] - want[1]).max()
        print(f"tile kernel, fast vs exact after 25 steps {kw}: max |du| {eu:.3e}, max |dv| {ev:.3e}")
        assert 0 < eu <= 1e-14 and ev <= 1e-14


def test_fast_flavour_spiral_traces_and_tip(oracle, yh):
    """2000 default-mode steps of a rotating spiral at 512^2 (C1 geometry): electrode traces and the tip
    of the FAST run against the EXACT run."""
    nx = ny = 512
    pe = yh.default_params(nx, ny, timeIntOrder=1, lap4=0)
    sim = yh.Sim(pe)
    sim.cross_field_ic()
    sim.run(12000, tb_steps=4)
    u0, v0 = (a[0].copy() for a in sim.get_state())
    sim.close()
    p = yh.default_params(nx, ny)
    probe = np.ix_(np.arange(40, ny, 108), np.arange(40, nx, 108))
    out = {}
    for flavour in (0, 1):
        assert yh.lib().yh_set_arithmetic(flavour) == 0
        uA, vA = dev(u0), dev(v0)
        uB, vB = torch.zeros_like(uA), torch.zeros_like(vA)
        tr = []
        ru, rv = uA, vA
        for seg in range(20):
            ou, ov = (uB, vB) if ru is uA else (uA, vA)
            ru, rv = host.rd_advance(p, 100, ru, rv, ou, ov, flags=host.RD_INPUT_CANONICAL if seg else 0)
            tr.append(ru.cpu().numpy()[probe].copy())
        prev = ru.clone()
        ou, ov = (uB, vB) if ru is uA else (uA, vA)
        ru, rv = host.rd_advance(p, 1, ru, rv, ou, ov, flags=host.RD_INPUT_CANONICAL)
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        vec = torch.zeros(65536 * 20, dtype=torch.uint8, device="cuda")
        host.tip_track(p, prev, ru, cnt, vec, t=0.0, capacity=65536)
        out[flavour] = (np.stack(tr), host.tips_to_numpy(cnt, vec), ru.cpu().numpy())
    assert yh.lib().yh_set_arithmetic(0) == 0
    d_tr = np.abs(out[0][0] - out[1][0]).max()
    d_u = np.abs(out[0][2] - out[1][2]).max()
    # The tip list also holds the spurious roots the shipped closed form reports wherever the sheet keeps an
    # exact bit-level symmetry of the initial condition; rounding differences move those.  The
    # physical tip -- the last entry, the one that centres the integration disc -- is what is compared.
    te, tf = out[0][1], out[1][1]
    assert len(te) > 0 and len(tf) > 0
    d_tip = max(abs(float(te[-1]["x"]) - float(tf[-1]["x"])), abs(float(te[-1]["y"]) - float(tf[-1]["y"])))
    print(f"fast vs exact over 2000 steps: electrodes {d_tr:.3e}, field {d_u:.3e}, spiral tip {d_tip:.3e} cell "
          f"({len(te)} / {len(tf)} list entries)")
    assert d_tr <= 1e-12 and d_u <= 1e-12 and d_tip <= 1e-4   # measured 1.0e-15, 4.6e-15, 0
    assert np.ptp(out[0][0], axis=0).max() > 0.02, "the spiral must move under the electrodes"
