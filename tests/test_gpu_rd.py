"""GPU parity of the reaction-diffusion step: the sm_100a kernels (through the C ABI) against
the plain-C oracle -- BITWISE in every mode (both sides compiled without FMA contraction) --
and against the reference's own kernels (oracle/_ref, --fmad=false) bitwise in its race-free
modes.  Tolerances, where any, are written in the test."""
import ctypes as C
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from tests import oracle_lib  # noqa: E402
from yolohtli_b200 import host, synth  # noqa: E402


@pytest.fixture(params=["stream", "tile", "quad"], autouse=True)
def rd_path(request):
    """Every test of this module runs three times: forced through the streaming kernels with a column
    pair per thread (rd_fast.cu / rd_rk.cu), through the shared-memory tile kernels (rd_tile.cu) where
    they apply, and through the four-columns-per-thread streaming Euler kernel (rd_quad.cu), which
    otherwise only serves sheets of 2 Mi cells and more."""
    import os
    if request.param == "quad" and request.node.originalname not in QUAD_TESTS:
        pytest.skip("no Euler kernel on this test's path")
    os.environ["YH_RD_PATH"] = "stream" if request.param == "quad" else request.param
    os.environ["YH_EULER_KERNEL"] = "quad" if request.param == "quad" else "pair"
    yield request.param
    os.environ.pop("YH_RD_PATH", None)
    os.environ.pop("YH_EULER_KERNEL", None)


QUAD_TESTS = {"test_every_mode_bitwise_vs_oracle", "test_temporal_blocking_is_bitwise_invariant",
              "test_masked_temporal_blocking_is_bitwise_invariant", "test_fast_path_gate_diff_off_and_negative_zero",
              "test_spiral_10k_steps_512", "test_slab_decomposition_is_bitwise_invariant",
              "test_bitwise_vs_reference_kernels_race_free_modes", "test_full_size_16384_sheet",
              "test_graph_replay_of_the_step_loop_is_bitwise_invariant"}


def dev(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a)).to("cuda", dtype=dtype).contiguous()


def rand_fields(nx, ny, seed=0):
    rng = np.random.default_rng(seed)
    return rng.uniform(-0.1, 1.1, (ny, nx)), rng.uniform(0.0, 1.0, (ny, nx))


def gpu_step(p, u, v, solid=None, velTan=False, **kw):
    du, dv = dev(u), dev(v)
    uo, vo = torch.empty_like(du), torch.empty_like(dv)
    vt = (torch.zeros_like(du), torch.zeros_like(du)) if velTan else None
    ds = dev(solid, torch.uint8) if solid is not None else None
    host.rd_step(p, du, dv, uo, vo, velTan=vt, solid=ds, **kw)
    torch.cuda.synchronize()
    out = [uo.cpu().numpy(), vo.cpu().numpy()]
    if velTan:
        out += [vt[0].cpu().numpy(), vt[1].cpu().numpy()]
    return out


def gpu_advance(p, n, u, v, tb=0, solid=None, **kw):
    uA, vA = dev(u), dev(v)
    uB, vB = torch.zeros_like(uA), torch.zeros_like(vA)
    ds = dev(solid, torch.uint8) if solid is not None else None
    ru, rv = host.rd_advance(p, n, uA, vA, uB, vB, tb_steps=tb, solid=ds, **kw)
    torch.cuda.synchronize()
    return ru.cpu().numpy(), rv.cpu().numpy()


MODES = []
for order, lap4, neu, so, gd, an in itertools.product((1, 2, 4), (0, 4), (1, 0), (0, 1), (1, 0), (0, 1)):
    MODES.append(dict(timeIntOrder=order, lap4=lap4, neumannBC=neu, solidSwitch=so, gateDiff=gd, anisotropy=an))


@pytest.mark.parametrize("size", [(48, 40), (50, 30), (130, 67)])
def test_every_mode_bitwise_vs_oracle(oracle, size):
    nx, ny = size
    u, v = rand_fields(nx, ny, 11)
    solid = (np.random.default_rng(5).uniform(size=(ny, nx)) > 0.2).astype(np.uint8)
    for m in MODES:
        p = oracle.params_default(nx, ny, **m)
        if m["anisotropy"]:
            oracle.l.yho_params_derive(C.byref(p), C.c_double(0.001), C.c_double(0.0004), C.c_double(0.0002))
        want = oracle.rd_step(p, u, v, solid=solid, velTan=True, stim_mouse=True, point=(nx // 3, ny // 2))
        got = gpu_step(p, u, v, solid=solid, velTan=True, stim_mouse=True, point=(nx // 3, ny // 2))
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), m
        if m["gateDiff"]:
            assert np.array_equal(got[2].ravel(), want[2].ravel()) and np.array_equal(got[3].ravel(), want[3].ravel()), m
        if m["solidSwitch"]:
            assert (got[0][solid == 0] == 0.0).all() and not np.signbit(got[0][solid == 0]).any()


@pytest.mark.parametrize("nx,ny", [(64, 48), (256, 96), (500, 300), (1030, 130), (2048, 64)])
@pytest.mark.parametrize("tb", [1, 2, 4])
def test_temporal_blocking_is_bitwise_invariant(oracle, nx, ny, tb):
    """T steps per HBM pass == T single-step launches == T oracle steps, bit for bit."""
    p = oracle.params_default(nx, ny, timeIntOrder=1, lap4=0)
    u, v = rand_fields(nx, ny, 21)
    n = 9
    want = oracle.rd_advance(p, n, u, v, stim_mouse=True, point=(nx // 2, ny // 3))
    got = gpu_advance(p, n, u, v, tb=tb, stim_mouse=True, point=(nx // 2, ny // 3))
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


@pytest.mark.parametrize("nx,ny", [(64, 48), (300, 200), (1024, 1024)])
@pytest.mark.parametrize("tb", [1, 2, 4])
def test_masked_temporal_blocking_is_bitwise_invariant(oracle, yh, nx, ny, tb):
    """Obstacle masks on the temporally blocked Euler kernel (C2: 1024^2 + holes): T steps per HBM
    pass == the oracle's single steps, bit for bit; non-tissue cells are exactly +0.0; lap4 is
    ignored in the mask branch (reactionDiffusion.cu:154-184).  Same through the headless driver."""
    u, v = rand_fields(nx, ny, 23)
    masks = [(np.random.default_rng(8).uniform(size=(ny, nx)) > 0.25).astype(np.uint8)]
    if nx == ny:
        masks.append(synth.hole_mask(nx, seed=4))
    n = 9
    for mask, gd in itertools.product(masks, (1, 0)):
        p = oracle.params_default(nx, ny, timeIntOrder=1, solidSwitch=1, gateDiff=gd)   # lap4 = 4 (default)
        want = oracle.rd_advance(p, n, u, v, solid=mask, stim_mouse=True, point=(nx // 2, ny // 3))
        got = gpu_advance(p, n, u, v, tb=tb, solid=mask, stim_mouse=True, point=(nx // 2, ny // 3))
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), (gd, mask.mean())
        assert (got[0][mask == 0] == 0.0).all() and not np.signbit(got[0][mask == 0]).any()
    if nx <= 300:
        p = oracle.params_default(nx, ny, timeIntOrder=1, lap4=0, solidSwitch=1)
        sim = yh.Sim(p, n_sims=2)
        sim.set_solid(masks[0])
        sim.set_state(np.stack([u, v]), np.stack([v, u]))
        sim.run(n, tb_steps=tb)
        su, sv = sim.get_state()
        sim.close()
        w0 = oracle.rd_advance(p, n, u, v, solid=masks[0])
        w1 = oracle.rd_advance(p, n, v, u, solid=masks[0])
        assert np.array_equal(su[0], w0[0]) and np.array_equal(sv[0], w0[1])
        assert np.array_equal(su[1], w1[0]) and np.array_equal(sv[1], w1[1])


@pytest.mark.parametrize("mode", ["euler", "euler_holes", "rk4lap4", "rk2"])
def test_graph_replay_of_the_step_loop_is_bitwise_invariant(oracle, yh, mode, monkeypatch):
    """Small sheets replay the step loop from a CUDA graph (64 time steps per graph, abi.cu): same
    bits as plain launches and as the oracle, on a cache miss, on a cache hit, with the result
    landing in either buffer pair, and through the headless driver (tips need the final
    single-step pass to stay last)."""
    nx, ny = 96, 80
    kw = dict(euler=dict(timeIntOrder=1, lap4=0), euler_holes=dict(timeIntOrder=1, lap4=0, solidSwitch=1),
              rk4lap4=dict(), rk2=dict(timeIntOrder=2, lap4=0))[mode]
    p = oracle.params_default(nx, ny, **kw)
    u, v = synth.cross_field_ic(nx, ny)
    u = u + 0.0
    u[40:50, 30:40] = -0.0                       # raw input: the first chunk must run un-graphed
    mask = (np.random.default_rng(2).uniform(size=(ny, nx)) > 0.1).astype(np.uint8) if "holes" in mode else None
    for n in (331, 256):                         # odd tail (T = 2, 1 at the end) and whole chunks
        want = oracle.rd_advance(p, n, u, v, solid=mask)
        monkeypatch.setenv("YH_GRAPHS", "0")
        plain = gpu_advance(p, n, u, v, tb=4, solid=mask)
        monkeypatch.setenv("YH_GRAPHS", "1")
        for rep in range(2):                     # miss, then hit (fresh tensors may or may not reuse addresses)
            got = gpu_advance(p, n, u, v, tb=4, solid=mask)
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), (n, rep)
            assert np.array_equal(got[0], plain[0]) and np.array_equal(got[1], plain[1])
    # same device buffers twice: the second call is a guaranteed cache hit; state continues
    uA, vA = dev(u), dev(v)
    uB, vB = torch.zeros_like(uA), torch.zeros_like(vA)
    ds = dev(mask, torch.uint8) if mask is not None else None
    ru, rv = host.rd_advance(p, 256, uA, vA, uB, vB, tb_steps=4, solid=ds)
    assert ru is uA
    ru, rv = host.rd_advance(p, 257, uA, vA, uB, vB, tb_steps=4, solid=ds, flags=host.RD_INPUT_CANONICAL)
    torch.cuda.synchronize()
    want = oracle.rd_advance(p, 513, u, v, solid=mask)
    assert ru is uB and np.array_equal(ru.cpu().numpy().reshape(ny, nx), want[0])
    assert np.array_equal(rv.cpu().numpy().reshape(ny, nx), want[1])
    if mask is None:
        sim = yh.Sim(p)
        sim.set_state(u, v)
        sim.run(321, tb_steps=4)
        su, sv = sim.get_state()
        w = oracle.rd_advance(p, 321, u, v)
        assert np.array_equal(su[0], w[0]) and np.array_equal(sv[0], w[1])
        prev = oracle.rd_advance(p, 320, u, v)
        assert sim.tips().tobytes() == oracle.tip_track(p, prev[0], w[0], t=p.dt * 321).tobytes()
        sim.close()


def test_fast_path_gate_diff_off_and_negative_zero(oracle):
    p = oracle.params_default(96, 80, timeIntOrder=1, lap4=0, gateDiff=0)
    u, v = rand_fields(96, 80, 3)
    u[10:20, 10:30] = 0.0
    v[10:20, 10:30] = 0.0
    u[30:40, 50:70] = -0.0   # raw user data may hold -0.0: the first pass forms u0 + 0.0 literally
    v[35:45, 40:60] = -0.0
    want = oracle.rd_advance(p, 6, u, v)
    got = gpu_advance(p, 6, u, v, tb=4)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert np.array_equal(np.signbit(got[0]), np.signbit(want[0]))


_ORACLE_RUNS = {}


def oracle_run_cached(oracle, key, p, n, u, v):
    """The long oracle runs are shared by the kernel-path variants of a test (same inputs)."""
    if key not in _ORACLE_RUNS:
        _ORACLE_RUNS[key] = oracle.rd_advance(p, n, u, v)
    return _ORACLE_RUNS[key]


def test_spiral_10k_steps_512(oracle):
    """BASELINE configs[0] (C1) as written: 512^2 cross-field spiral, headless, 10 000 fixed steps,
    Euler + 5-point, bitwise vs the plain-C host loop."""
    p = oracle.params_default(512, 512, timeIntOrder=1, lap4=0)
    u, v = synth.cross_field_ic(512, 512)
    want = oracle_run_cached(oracle, "c1_euler_10k", p, 10000, u, v)
    got = gpu_advance(p, 10000, u, v, tb=4)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert 0.05 < want[0].mean() < 0.9   # a wave is actually propagating


def test_spiral_4k_steps_512_default_mode(oracle, rd_path):
    """C1 in the reference's DEFAULT mode (RK4 + 4th-order Laplacian + gateDiff, saveFiles.cu:124-132):
    4 000 steps (the CPU side of 10 000 takes minutes) bitwise vs the plain-C host loop with
    synchronous stages."""
    p = oracle.params_default(512, 512)
    u, v = synth.cross_field_ic(512, 512)
    want = oracle_run_cached(oracle, "c1_default_4k", p, 4000, u, v)
    got = gpu_advance(p, 4000, u, v)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert 0.05 < want[0].mean() < 0.9


def test_default_mode_rk4_lap4_vs_oracle_and_holes(oracle):
    # C1 default mode and C2 (1024^2-style mask, scaled down): bitwise vs the synchronous oracle
    p = oracle.params_default(256, 256)
    assert p.timeIntOrder == 4 and p.lap4 == 4
    u, v = synth.fibrillation_ic(256, 256, patch=64)
    want = oracle.rd_advance(p, 20, u, v)
    got = gpu_advance(p, 20, u, v)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    mask = synth.hole_mask(256, seed=3)
    p = oracle.params_default(256, 256, solidSwitch=1)
    want = oracle.rd_advance(p, 20, u * mask, v * mask, solid=mask)
    got = gpu_advance(p, 20, u * mask, v * mask, solid=mask)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert (got[0][mask == 0] == 0.0).all()


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("masked", [False, True])
def test_slab_decomposition_is_bitwise_invariant(oracle, world, masked):
    """N row slabs with ghost rows (emulated on one GPU, halos copied between slab buffers)
    == the single-domain run, bit for bit (multi-GPU invariance); with and
    without obstacle masks (each slab passes its own rows of the mask, ghost rows included)."""
    from yolohtli_b200.slab import SlabLayout
    nx, ny, H, nsteps = 256, 211, 4, 12
    pg = oracle.params_default(nx, ny, timeIntOrder=1, lap4=0, solidSwitch=int(masked))
    u, v = rand_fields(nx, ny, 31)
    mask = (np.random.default_rng(12).uniform(size=(ny, nx)) > 0.2).astype(np.uint8) if masked else None
    want = oracle.rd_advance(pg, nsteps, u, v, solid=mask, stim_mouse=True, point=(100, 105))
    lays = [SlabLayout(ny, world, r, H) for r in range(world)]
    dmask = [dev(mask[l.g0:l.g1], torch.uint8) if masked else None for l in lays]
    mflags = 0
    if masked and world == 2:   # fixed mask: neighbourhood patterns derived once per slab (yh_rd_mask_patterns)
        dmask = [host.rd_mask_patterns(l.local_params(pg), m, torch.empty_like(m)) for l, m in zip(lays, dmask)]
        mflags = host.RD_SOLID_IS_PATTERNS
    bufs = []
    for l in lays:
        a = [dev(u[l.g0:l.g1]), dev(v[l.g0:l.g1])]
        bufs.append((a, [torch.zeros_like(a[0]), torch.zeros_like(a[1])]))
    cur = [b[0] for b in bufs]
    oth = [b[1] for b in bufs]
    for it in range(nsteps // H):
        if it > 0:   # halo exchange: owned edge rows -> neighbours' ghosts
            for r, l in enumerate(lays):
                if l.down is not None:
                    d = lays[l.down]
                    for f in range(2):
                        cur[l.down][f][d.own_lo - H:d.own_lo] = cur[r][f][l.own_hi - H:l.own_hi]
                        cur[r][f][l.own_hi:l.own_hi + H] = cur[l.down][f][d.own_lo:d.own_lo + H]
        for r, l in enumerate(lays):
            p = l.local_params(pg)
            ru, rv = host.rd_advance(p, H, cur[r][0], cur[r][1], oth[r][0], oth[r][1], tb_steps=4,
                                     rows=(l.own_lo, l.own_hi), stim_mouse=True, point=(100, 105),
                                     solid=dmask[r], flags=mflags)
            if ru is oth[r][0]:
                cur[r], oth[r] = oth[r], cur[r]
    torch.cuda.synchronize()
    got_u = np.concatenate([cur[r][0][l.own_lo:l.own_hi].cpu().numpy() for r, l in enumerate(lays)])
    got_v = np.concatenate([cur[r][1][l.own_lo:l.own_hi].cpu().numpy() for r, l in enumerate(lays)])
    assert np.array_equal(got_u, want[0]) and np.array_equal(got_v, want[1])


def test_slab_generic_rk4_rows(oracle):
    # multi-stage path on a slab: ghost depth = stages
    from yolohtli_b200.slab import SlabLayout
    nx, ny = 64, 50
    pg = oracle.params_default(nx, ny)   # RK4 + lap4
    u, v = rand_fields(nx, ny, 41)
    want = oracle.rd_step(pg, u, v)
    for r in range(2):
        l = SlabLayout(ny, 2, r, 4)
        p = l.local_params(pg)
        got = gpu_step(p, u[l.g0:l.g1], v[l.g0:l.g1], rows=(l.own_lo, l.own_hi))
        assert np.array_equal(got[0][l.own_lo:l.own_hi], want[0][l.j0:l.j1])
        assert np.array_equal(got[1][l.own_lo:l.own_hi], want[1][l.j0:l.j1])


@pytest.mark.parametrize("order,nsteps", [(1, 3), (4, 1), (2, 2)])
def test_anisotropic_slab_rows_bitwise(oracle, order, nsteps):
    """Anisotropy + no-flux boundaries on a ROW RANGE: the corner corrections at x = 0 / x = nx-1 read rows
    j +- 2 (reactionDiffusion.cu:290-304), so a stage consumes two ghost rows; the slab result must equal the
    whole-sheet result bit for bit, and a slab with too few ghost rows must be refused, not computed wrong."""
    nx, ny = 96, 120
    p = oracle.params_default(nx, ny, timeIntOrder=order, lap4=0, anisotropy=1)
    p.rxy, p.rbx, p.rby = 0.013, 0.21, 0.19          # a non-trivial diffusion tensor
    u, v = rand_fields(nx, ny, seed=9)
    wu, wv = oracle.rd_advance(p, nsteps, u, v)
    lo, hi = 40, 90
    H = 2 * order * nsteps                            # two rows per stage and step
    q = p.copy()
    q.ny, q.ny_global, q.jg0 = hi - lo + 2 * H, ny, lo - H
    gu, gv = gpu_advance(q, nsteps, u[lo - H:hi + H], v[lo - H:hi + H], rows=(H, H + hi - lo))
    assert np.array_equal(gu[H:H + hi - lo], wu[lo:hi]) and np.array_equal(gv[H:H + hi - lo], wv[lo:hi])
    # half the ghost rows: refused
    Hs = order * nsteps
    q.ny, q.jg0 = hi - lo + 2 * Hs, lo - Hs
    with pytest.raises(Exception):
        gpu_advance(q, nsteps, u[lo - Hs:hi + Hs], v[lo - Hs:hi + Hs], rows=(Hs, Hs + hi - lo))


@pytest.mark.skipif(not oracle_lib.have_reference(), reason="oracle/_ref not built")
def test_bitwise_vs_reference_kernels_race_free_modes(oracle):
    """T1 tier: Euler + lap4=0, every boundary/mask branch the reference defines, against the
    reference's OWN kernels (sm_100, --fmad=false): bit-identical after 25 steps."""
    ref = oracle_lib.Reference(nofma=True)
    nx, ny = 144, 112
    u, v = rand_fields(nx, ny, 51)
    solid = (np.random.default_rng(6).uniform(size=(ny, nx)) > 0.1).astype(np.uint8)
    for neu, so, gd in itertools.product((1, 0), (0, 1), (1, 0)):
        if (not neu) and so:
            continue   # reference indexes out of bounds at the edges in this branch
        p = oracle.params_default(nx, ny, timeIntOrder=1, lap4=0, neumannBC=neu, solidSwitch=so, gateDiff=gd)
        ref.init(p)
        ru, rv, _ = ref.rd_run(u, v, 25, solid=solid, stim_mouse=True, point=(70, 50))
        gu, gv = gpu_advance(p, 25, u, v, tb=4, solid=solid if so else None, stim_mouse=True, point=(70, 50))
        assert np.array_equal(gu, ru) and np.array_equal(gv, rv), (neu, so, gd)
    # with the reference's DEFAULT flags (FMA contraction on) the fields agree to 1e-12
    ref2 = oracle_lib.Reference(nofma=False)
    p = oracle.params_default(nx, ny, timeIntOrder=1, lap4=0)
    ref2.init(p)
    ru, rv, _ = ref2.rd_run(u, v, 25)
    gu, gv = gpu_advance(p, 25, u, v, tb=4)
    assert np.abs(gu - ru).max() < 1e-12 and np.abs(gv - rv).max() < 1e-12


@pytest.mark.skipif(not oracle_lib.have_reference(), reason="oracle/_ref not built")
def test_racy_reference_modes_within_tolerance(oracle):
    """T2 tier: the reference's default RK4 + lap4 kernel races on g_in / J
    (reactionDiffusion.cu:117,219); against our synchronous stages the smooth spiral fields
    agree to 2e-3 after 200 steps (the reference's own run-to-run spread is of that order)."""
    ref = oracle_lib.Reference(nofma=False)
    p = oracle.params_default(256, 256)
    u, v = synth.cross_field_ic(256, 256)
    ref.init(p)
    a = ref.rd_run(u, v, 200)
    b = ref.rd_run(u, v, 200)
    spread = max(np.abs(a[0] - b[0]).max(), 1e-16)
    gu, gv = gpu_advance(p, 200, u, v)
    err = np.abs(gu - a[0]).max()
    print(f"racy reference: run-to-run spread {spread:.3e}, ours vs reference {err:.3e}")
    assert err < 2e-3


@pytest.mark.skipif(not oracle_lib.have_reference(), reason="oracle/_ref not built")
def test_default_mode_trace_and_tip_trajectory_vs_reference(oracle, yh, rd_path):
    """North-star acceptance for the reference's DEFAULT mode (RK4 + 4th-order Laplacian).  The
    reference kernel races in this mode (reactionDiffusion.cu:117,219) and is not reproducible run
    to run, so the bound is calibrated on the reference itself (tier T2): over 1600 steps
    (32 ms) of a rotating spiral, sampled every sampleIt = 100 steps, our synchronous-stage path must
    stay within max(0.5 cell, 3x the spread of FIVE reference runs) in tip position (median <= 0.4
    cell) and within max(1e-3, 3x spread) in the voltage of a 5 x 5 grid of electrodes.
    The tip LIST is checked separately and exactly: the reference's own tip kernel (--fmad=false
    build) applied to OUR fields returns our list bit for bit -- including the spurious roots the
    shipped closed form (no residual check) reports wherever the sheet still has the x-only
    symmetry of the initial condition exactly, which our arithmetic preserves and the reference's
    races break at 1e-19."""
    if rd_path != "stream":
        pytest.skip("one path is enough for this statistical tier")
    nx = ny = 256
    nseg = 16   # beyond ~2000 steps the reference's tip starts splitting into 2-3 noisy crossings
    pe = oracle.params_default(nx, ny, timeIntOrder=1, lap4=0)
    sim = yh.Sim(pe)
    sim.cross_field_ic()
    sim.run(9000, tb_steps=4)           # let the cross-field IC curl into a spiral (bitwise-pinned Euler path)
    u0, v0 = (a[0] for a in sim.get_state())
    sim.close()
    p = oracle.params_default(nx, ny)   # default: RK4 + lap4
    ref = oracle_lib.Reference(nofma=False)
    ref.init(p)
    probe = np.ix_(np.arange(24, ny, 52), np.arange(24, nx, 52))   # 5 x 5 electrodes
    r_path, r_trace, r_all = [], [], {}
    for run in range(5):
        ru, rv = u0, v0
        path, trace = [], []
        for seg in range(nseg):
            ru, rv, _ = ref.rd_run(ru, rv, 99)
            pu_r = ru
            ru, rv, _ = ref.rd_run(ru, rv, 1)
            t_r = ref.tip(ru, pu_r, t=0.0, algorithm=1)      # (present, past), main.cu:963
            assert 1 <= len(t_r) <= 5, (run, seg, t_r)       # one spiral; the racy fields can split its tip
            assert np.ptp(t_r["x"]) < 4 and np.ptp(t_r["y"]) < 4, (run, seg, t_r)
            path.append((float(t_r["x"].mean()), float(t_r["y"].mean())))
            r_all.setdefault(seg, []).extend(zip(t_r["x"].tolist(), t_r["y"].tolist()))
            trace.append(ru[probe])
        r_path.append(path); r_trace.append(trace)
    r_path, r_trace = np.array(r_path), np.array(r_trace).reshape(5, nseg, -1)    # [run, seg, 2], [run, seg, electrode]
    refn = oracle_lib.Reference(nofma=True)
    refn.init(p)
    ours = yh.Sim(p)
    ours.set_state(u0, v0)
    o_path, o_trace = [], []
    for seg in range(nseg):
        ours.run(99, tb_steps=1)
        pu_o = ours.get_state()[0][0].copy()
        ours.run(1, tb_steps=1)
        ou = ours.get_state()[0][0]
        t_o = ours.tips()
        t_x = refn.tip(ou, pu_o, t=p.dt * ours.count, algorithm=1)
        key = ("y", "x")
        assert np.array_equal(np.sort(t_o, order=key), np.sort(t_x, order=key)), seg
        centre = r_path[:, seg].mean(axis=0)
        d = np.hypot(t_o["x"] - centre[0], t_o["y"] - centre[1])
        o_path.append((float(t_o["x"][d.argmin()]), float(t_o["y"][d.argmin()])))
        o_trace.append(ou[probe])
    ours.close()
    o_path, o_trace = np.array(o_path), np.array(o_trace).reshape(nseg, -1)
    spread_tip = np.array([max(np.hypot(*(r_path[a, k] - r_path[b, k])) for a in range(5) for b in range(5))
                           for k in range(nseg)])
    spread_u = r_trace.max(axis=0) - r_trace.min(axis=0)
    dev_tip = np.array([min(np.hypot(o_path[k][0] - x, o_path[k][1] - y) for x, y in r_all[k]) for k in range(nseg)])
    dev_u = np.abs(o_trace[None] - r_trace).min(axis=0)
    print("reference run-to-run tip spread (cells):", np.round(spread_tip, 3))
    print("ours - nearest reference run      (cells):", np.round(dev_tip, 3))
    print("reference trace spread (max over electrodes):", np.round(spread_u.max(axis=1), 6))
    print("ours - nearest reference trace (max over electrodes):", np.round(dev_u.max(axis=1), 6))
    # measured on B200: ours sits a steady 0.10-0.33 cell from the reference path.  The racy stage
    # reads make the reference a slightly different scheme whose outcome depends on the state of
    # the GPU it runs on (0.15 cell alone, 0.3 cell inside the full suite) while repeated runs in
    # one process differ by only 0.01 cell; once its tip starts splitting the spread is 0.5-1.4.
    assert (dev_tip <= np.maximum(0.5, 3.0 * spread_tip)).all() and np.median(dev_tip) <= 0.4
    assert (dev_u <= np.maximum(1e-3, 3.0 * spread_u)).all()   # measured: <= 2.7e-4 (reference spread 5e-5)
    assert (np.ptp(r_trace[0], axis=0) > 0.02).sum() >= 2, "some electrodes must see the voltage move"


def test_full_size_16384_sheet(oracle, rd_path):
    """BASELINE configs[3] at FULL size (16384 x 16384, 8 GiB of state): 4 time steps in ONE
    temporally-blocked pass == 4 single-step passes == the plain-C oracle, bit for bit, plus the
    size-independent properties: untouched input, row-band locality (a band recomputed alone with
    ghost rows reproduces the same bits), and a checksum of checksums over row blocks."""
    import psutil
    if rd_path == "tile":
        pytest.skip("full size is the streaming kernels' regime")
    free, total = torch.cuda.mem_get_info()
    if free < 24 << 30 or psutil.virtual_memory().available < 40 << 30:
        pytest.skip("needs 24 GiB of HBM and 40 GiB of host memory")
    n = 16384
    p = oracle.params_default(n, n, scale_L=True, timeIntOrder=1, lap4=0)
    u0, v0 = synth.fibrillation_ic(n, n)
    uA, vA = torch.as_tensor(u0).cuda(), torch.as_tensor(v0).cuda()
    uB, vB = torch.empty_like(uA), torch.empty_like(vA)
    uC, vC = torch.empty_like(uA), torch.empty_like(vA)
    r4u, r4v = host.rd_advance(p, 4, uA, vA, uB, vB, tb_steps=4)          # one pass, T = 4
    assert r4u is uB
    assert torch.equal(uA.cpu(), torch.as_tensor(u0))                      # input never written
    uD, vD = uA.clone(), vA.clone()
    r1u, r1v = host.rd_advance(p, 4, uD, vD, uC, vC, tb_steps=1)          # four passes, T = 1
    assert torch.equal(r4u, r1u) and torch.equal(r4v, r1v)
    # row-band locality: rows [6000, 6512) recomputed from a slab with 4 ghost rows per side
    from yolohtli_b200.slab import SlabLayout
    lo, hi, H = 6000, 6512, 4
    q = p.copy()
    q.ny, q.ny_global, q.jg0 = hi - lo + 2 * H, n, lo - H
    su, sv = uA[lo - H:hi + H].contiguous(), vA[lo - H:hi + H].contiguous()
    tu, tv = torch.empty_like(su), torch.empty_like(sv)
    bu, bv = host.rd_advance(q, 4, su, sv, tu, tv, tb_steps=4, rows=(H, H + hi - lo))
    assert torch.equal(bu[H:H + hi - lo], r4u[lo:hi]) and torch.equal(bv[H:H + hi - lo], r4v[lo:hi])
    # the oracle at full size (a few seconds on the host cores)
    wu, wv = oracle.rd_advance(p, 4, u0, v0)
    gu = r4u.cpu().numpy()
    assert np.array_equal(gu, wu) and np.array_equal(r4v.cpu().numpy(), wv)
    blocks = gu.reshape(64, 256, n).sum(axis=(1, 2))
    assert blocks.sum() == wu.reshape(64, 256, n).sum(axis=(1, 2)).sum()
