"""GPU: the marching tile kernel of the small-sheet default mode (csrc/rd_tile_march.cu).

EXACT flavour: bit-identical to the plain-C oracle for RK2 / RK4, with / without the 4th-order Laplacian,
default and non-default model constants, velTan output, row ranges of a slab (mirrors only at the GLOBAL
edges), sheets narrower / lower than one tile, and under forced tilings whose strips and bands cut the
sheet at odd places.  FAST flavour: within 1e-14 of the exact result after 25 steps (measured 4e-16)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from yolohtli_b200 import host, synth  # noqa: E402


@pytest.fixture(autouse=True)
def march_kernel():
    os.environ["YH_RD_PATH"] = "tile"
    os.environ["YH_TILE_RK"] = "march"
    yield
    for k in ("YH_RD_PATH", "YH_TILE_RK", "YH_MARCH_TILING"):
        os.environ.pop(k, None)


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def advance(p, n, u, v, rows=None, solid=None):
    uA, vA = dev(u), dev(v)
    uB, vB = torch.zeros_like(uA), torch.zeros_like(vA)
    ds = torch.as_tensor(np.ascontiguousarray(solid, dtype=np.uint8)).cuda() if solid is not None else None
    ru, rv = host.rd_advance(p, n, uA, vA, uB, vB, rows=rows, solid=ds)
    torch.cuda.synchronize()
    return ru.cpu().numpy(), rv.cpu().numpy()


def fields(nx, ny, seed=0):
    rng = np.random.default_rng(seed)
    return rng.uniform(-0.1, 1.1, (ny, nx)), rng.uniform(0.0, 1.0, (ny, nx))


KW = [dict(), dict(lap4=0), dict(timeIntOrder=2), dict(timeIntOrder=2, lap4=0),
      dict(mu=1.1, delta=0.9, gamma=0.05, theta=0.01, tc=0.9), dict(lap4=0, alpha=0.15, eps=0.01)]


@pytest.mark.parametrize("nx,ny", [(8, 8), (48, 40), (50, 30), (130, 67), (256, 96), (500, 131), (512, 512), (640, 37)])
@pytest.mark.parametrize("kw", KW)
def test_march_exact_bitwise_vs_oracle(oracle, nx, ny, kw):
    p = oracle.params_default(nx, ny, **kw)
    u, v = fields(nx, ny, seed=nx + ny)
    u[1:5, 2:8] = -0.0          # raw user data may hold -0.0
    want = oracle.rd_advance(p, 3, u, v)
    got = advance(p, 3, u, v)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert np.array_equal(np.signbit(got[0]), np.signbit(want[0]))


@pytest.mark.parametrize("tiling", ["8,5", "22,9", "56,40", "30,17", "52,1"])
@pytest.mark.parametrize("kw", [dict(), dict(timeIntOrder=2, lap4=0)])
def test_march_forced_tilings(oracle, tiling, kw):
    """Strips and bands that do not divide the sheet, bands lower than the halo, one-row bands."""
    os.environ["YH_MARCH_TILING"] = tiling
    nx, ny = 118, 61
    p = oracle.params_default(nx, ny, **kw)
    u, v = fields(nx, ny, seed=3)
    want = oracle.rd_advance(p, 2, u, v)
    got = advance(p, 2, u, v)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


def test_march_veltan_and_row_ranges(oracle):
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
    gu, gv = advance(q, 1, u[lo - H:hi + H], v[lo - H:hi + H], rows=(H, H + hi - lo))
    assert np.array_equal(gu[H:H + hi - lo], wu[lo:hi]) and np.array_equal(gv[H:H + hi - lo], wv[lo:hi])
    # ... slabs that touch the top / the bottom of the sheet (no ghost rows there, no-flux mirror instead),
    # and slabs whose first owned row is closer to the global edge than the halo
    for lo, hi in [(0, 60), (70, ny), (2, 50), (80, ny - 1)]:
        g0, g1 = max(0, lo - H), min(ny, hi + H)
        q.ny, q.jg0 = g1 - g0, g0
        gu, gv = advance(q, 1, u[g0:g1], v[g0:g1], rows=(lo - g0, hi - g0))
        assert np.array_equal(gu[lo - g0:hi - g0], wu[lo:hi]) and np.array_equal(gv[lo - g0:hi - g0], wv[lo:hi]), (lo, hi)


def test_march_equals_first_tile_kernel_over_200_steps(oracle):
    """The two tile kernels are two schedules of the same expressions: 200 steps of the spiral, bit for bit
    (the run goes through the CUDA-graph replay of the step loop)."""
    nx = ny = 512
    p = oracle.params_default(nx, ny)
    u, v = synth.cross_field_ic(nx, ny)
    a = advance(p, 200, u, v)
    os.environ["YH_TILE_RK"] = "cell"
    b = advance(p, 200, u, v)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("kw", [dict(), dict(lap4=0), dict(timeIntOrder=2), dict(mu=1.1, delta=0.9, gamma=0.05, theta=0.01)])
def test_march_fast_flavour_25_steps_within_1e14(oracle, yh, kw):
    nx = ny = 512
    p = oracle.params_default(nx, ny, **kw)
    u, v = synth.cross_field_ic(nx, ny)
    u = u + 0.05 * np.sin(0.07 * np.arange(nx))[None, :] * np.cos(0.05 * np.arange(ny))[:, None]
    want = oracle.rd_advance(p, 25, u, v)
    assert yh.lib().yh_set_arithmetic(1) == 0
    try:
        got = advance(p, 25, u, v)
    finally:
        assert yh.lib().yh_set_arithmetic(0) == 0
    eu, ev = np.abs(got[0] - want[0]).max(), np.abs(got[1] - want[1]).max()
    print(f"march fast vs exact after 25 steps {kw}: max |du| {eu:.3e}, max |dv| {ev:.3e}")
    assert 0 < eu <= 1e-14 and ev <= 1e-14


def random_mask(nx, ny, seed):
    """Tissue with holes of every local shape: isolated cells, bars, blocks, holes on the sheet's edges."""
    rng = np.random.default_rng(seed)
    m = (rng.uniform(size=(ny, nx)) > 0.08).astype(np.uint8)
    for _ in range(6):
        j, i = rng.integers(0, ny), rng.integers(0, nx)
        m[j:j + rng.integers(1, 9), i:i + rng.integers(1, 17)] = 0
    m[0, :7] = 0
    m[-1, -5:] = 0
    m[ny // 2:, 0] = 0
    return m


@pytest.mark.parametrize("nx,ny", [(48, 40), (130, 67), (256, 96), (500, 131), (1024, 300)])
@pytest.mark.parametrize("kw", [dict(), dict(timeIntOrder=2), dict(mu=1.1, delta=0.9, gamma=0.05, theta=0.01, tc=0.9)])
def test_march_masks_bitwise_vs_oracle(oracle, nx, ny, kw):
    """Obstacle masks (C2) on the marching tiles: mask codes derived once per launch, kept in registers."""
    os.environ["YH_SOLID_RK"] = "march"
    try:
        p = oracle.params_default(nx, ny, solidSwitch=1, **kw)
        mask = random_mask(nx, ny, nx + 7 * ny)
        u, v = fields(nx, ny, seed=nx)
        u, v = u * mask, v * mask
        want = oracle.rd_advance(p, 3, u, v, solid=mask)
        got = advance(p, 3, u, v, solid=mask)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        assert (got[0][mask == 0] == 0.0).all() and not np.signbit(got[0][mask == 0]).any()
        os.environ["YH_SOLID_RK"] = "stream"
        other = advance(p, 3, u, v, solid=mask)
        assert np.array_equal(got[0], other[0]) and np.array_equal(got[1], other[1])
        # a slab of the same sheet: rows [lo, hi) with ghost rows, mask rows stored alike
        os.environ["YH_SOLID_RK"] = "march"
        if ny >= 60:
            H = 4
            wu, wv = oracle.rd_advance(p, 1, u, v, solid=mask)
            for lo, hi in [(0, 30), (20, ny - 10), (ny - 25, ny)]:
                g0, g1 = max(0, lo - H), min(ny, hi + H)
                q = p.copy()
                q.ny, q.ny_global, q.jg0 = g1 - g0, ny, g0
                gu, gv = advance(q, 1, u[g0:g1], v[g0:g1], rows=(lo - g0, hi - g0), solid=mask[g0:g1])
                assert np.array_equal(gu[lo - g0:hi - g0], wu[lo:hi]) and np.array_equal(gv[lo - g0:hi - g0], wv[lo:hi]), (lo, hi)
    finally:
        os.environ.pop("YH_SOLID_RK", None)
