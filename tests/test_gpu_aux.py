"""GPU parity of tip tracking, APD bookkeeping, probe, the symmetry-reduction kernels and the
headless driver against the plain-C oracle (bitwise) and the reference's own kernels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from tests import oracle_lib  # noqa: E402
from yolohtli_b200 import host, synth  # noqa: E402

TIP_DTYPE = oracle_lib.TIP_DTYPE


def dev(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a)).to("cuda", dtype=dtype).contiguous()


def fields(nx, ny, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(-0.1, 1.1, (ny, nx)), rng.uniform(0.0, 1.0, (ny, nx))


def wavy(nx, ny, a=0.31, b=0.27, ph=0.1):
    X, Y = np.meshgrid(np.arange(nx, dtype=float), np.arange(ny, dtype=float))
    return 0.7 + 0.3 * np.sin(a * X + ph) * np.cos(b * Y) + 0.05 * np.sin(0.11 * X * Y / nx)


def gpu_tips(p, past, present, t=0.0, alg=None, plot=False, capacity=host.TIPVECSIZE):
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    vec = torch.zeros(capacity * 20, dtype=torch.uint8, device="cuda")
    pl = torch.zeros(p.nx * p.ny, dtype=torch.uint8, device="cuda") if plot else None
    host.tip_track(p, dev(past), dev(present), cnt, vec, tip_plot=pl, t=t, algorithm=alg, capacity=capacity)
    torch.cuda.synchronize()
    out = host.tips_to_numpy(cnt, vec)
    return (out, pl.cpu().numpy()) if plot else out


@pytest.mark.parametrize("nx,ny", [(48, 40), (333, 257), (1024, 1024)])
@pytest.mark.parametrize("alg", [1, 2])
def test_tips_bitwise_and_ordered(oracle, nx, ny, alg):
    p = oracle.params_default(nx, ny, tipOffsetX=min(160, nx // 2 - 2), tipOffsetY=min(160, ny // 2 - 2))
    past, present = wavy(nx, ny), wavy(nx, ny, 0.29, 0.33, 0.7)
    want, wplot = oracle.tip_track(p, past, present, t=2.5, algorithm=alg, plot=True)
    for rep in range(3):   # ordered compaction: identical list every launch
        got, gplot = gpu_tips(p, past, present, t=2.5, alg=alg, plot=True)
        assert len(got) == len(want) and len(want) > 0
        assert got.tobytes() == want.tobytes()
        assert np.array_equal(gplot, wplot)


def test_tips_gradient_solid_and_flag_algorithm(oracle):
    nx = ny = 128
    past, present = wavy(nx, ny), wavy(nx, ny, 0.29, 0.33, 0.7)
    p = oracle.params_default(nx, ny, tipGrad=1, tipOffsetX=40, tipOffsetY=40)
    assert gpu_tips(p, past, present).tobytes() == oracle.tip_track(p, past, present).tobytes()
    p = oracle.params_default(nx, ny, solidSwitch=1, tipOffsetX=40, tipOffsetY=40)
    got, want = gpu_tips(p, past, present), oracle.tip_track(p, past, present)
    assert len(want) > 0 and got.tobytes() == want.tobytes()
    p = oracle.params_default(nx, ny)
    g, gp = gpu_tips(p, past, present, alg=3, plot=True)
    w, wp = oracle.tip_track(p, past, present, algorithm=3, plot=True)
    assert len(g) == 0 and np.array_equal(gp, wp) and wp.sum() > 0


def test_tips_empty_and_capacity(oracle):
    p = oracle.params_default(64, 64)
    z = np.zeros((64, 64))
    assert len(gpu_tips(p, z, z)) == 0
    past, present = wavy(64, 64), wavy(64, 64, 0.29, 0.33, 0.7)
    want = oracle.tip_track(p, past, present)
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    vec = torch.zeros(2 * 20 + 20, dtype=torch.uint8, device="cuda")
    host.tip_track(p, dev(past), dev(present), cnt, vec, capacity=2)
    torch.cuda.synchronize()
    assert int(cnt.item()) == len(want) > 2          # count keeps the true total
    assert vec[40:].sum().item() == 0                # nothing written past the capacity


@pytest.mark.skipif(not oracle_lib.have_reference(), reason="oracle/_ref not built")
def test_tips_vs_reference_kernels(oracle):
    ref = oracle_lib.Reference(nofma=True)
    nx = ny = 256
    p = oracle.params_default(nx, ny)
    ref.init(p)
    past, present = wavy(nx, ny), wavy(nx, ny, 0.29, 0.33, 0.7)
    r = ref.tip(present, past, t=1.0, algorithm=1)       # unordered (atomicAdd)
    g = gpu_tips(p, past, present, t=1.0, alg=1)
    assert len(r) == len(g) > 0                          # T0: tip count bit-exact
    key = lambda t: np.sort(np.stack([t["x"], t["y"], t["t"]], 1).view("f4,f4,f4").reshape(-1))
    assert np.array_equal(key(r), key(g))                # same records as a multiset


def test_sapd_sequence_bitwise(oracle):
    nx, ny = 96, 64
    p = oracle.params_default(nx, ny)
    X, Y = np.meshgrid(np.arange(nx, dtype=float), np.arange(ny, dtype=float))
    seq = [np.clip(0.5 + 0.6 * np.sin(0.2 * k + 0.05 * X + 0.07 * Y), -0.1, 1.1) for k in range(80)]
    area = synth.stim_area_square(nx, ny)
    for stimulate in (False, True):
        want = oracle.sapd_sequence(p, seq, count0=1, stimArea=area, stimulate=stimulate)
        n = nx * ny
        st = {k: torch.zeros(n, dtype=torch.float64, device="cuda") for k in ("APD1", "APD2", "sAPD", "dAPD", "back", "front")}
        first = torch.zeros(n, dtype=torch.uint8, device="cuda")
        dseq = [dev(a) for a in seq]
        for k in range(len(seq) - 1):
            host.sapd(p, 1 + k, dseq[k], dseq[k + 1], st["APD1"], st["APD2"], st["sAPD"], st["dAPD"],
                      st["back"], st["front"], first, stimArea=dev(area, torch.uint8), stimulate=stimulate)
        torch.cuda.synchronize()
        for k in ("APD1", "APD2", "sAPD", "back", "front") + (("dAPD",) if stimulate else ()):
            assert np.array_equal(st[k].cpu().numpy(), want[k]), k
        assert np.array_equal(first.cpu().numpy(), want["first"])
        assert want["APD1"].max() > 0 and want["APD2"].max() > 0


def test_probe(oracle):
    p = oracle.params_default(64, 48)
    u, v = fields(64, 48, 1)
    pt = torch.zeros(2, dtype=torch.float64, device="cuda")
    host.probe(p, dev(u), dev(v), pt, 17, 31)
    assert pt.cpu().tolist() == [u[31, 17], v[31, 17]]


@pytest.mark.parametrize("nx,ny,off", [(96, 96, 30), (512, 512, 160), (200, 140, 50)])
def test_sr_kernels_bitwise(oracle, nx, ny, off):
    p = oracle.params_default(nx, ny, reduce_sym=True, tipOffsetX=off, tipOffsetY=off,
                              tipx0=nx / 2 + 3.0, tipy0=ny / 2 - 5.0)
    u, v = fields(nx, ny, 2)
    vtu, vtv = fields(nx, ny, 3)
    c, phi = [0.13, -0.21, 0.04], [0.3, -0.1, 0.77]
    # Cxy
    wax, way = oracle.cxy_field(p, c, phi)
    ax, ay = torch.zeros(nx * ny, dtype=torch.float64, device="cuda"), torch.zeros(nx * ny, dtype=torch.float64, device="cuda")
    host.cxy_field(p, ax, ay, c, phi)
    assert np.array_equal(ax.cpu().numpy(), wax) and np.array_equal(ay.cpu().numpy(), way)
    # slice (both schemes) + trapz + fused
    du, dv, dvtu, dvtv = dev(u), dev(v), dev(vtu), dev(vtv)
    for scheme in (2, 1):
        ws, ws0 = oracle.slice(p, u, v, wax, way, scheme=scheme)
        s = [torch.full((nx * ny,), 7.0, dtype=torch.float64, device="cuda") for _ in range(6)]
        s0 = [torch.full((nx * ny,), 7.0, dtype=torch.float64, device="cuda") for _ in range(6)]
        host.slice_fields(p, du, dv, s, s0, ax, ay, scheme=scheme)
        for q in range(6):
            assert np.array_equal(s[q].cpu().numpy(), ws[q]), (scheme, q)
            assert np.array_equal(s0[q].cpu().numpy(), ws0[q]), (scheme, q)
        if scheme == 2:
            want = oracle.trapz(p, ws, ws0, vtu, vtv)
            got = host.trapz(p, s, s0, dvtu, dvtv)
            assert np.array_equal(got, want)
            fused = host.sr_integrals(p, du, dv, dvtu, dvtv, ax, ay)
            assert np.array_equal(fused, want)
    # disc centre from the device tip list
    tips = np.zeros(3, dtype=TIP_DTYPE)
    tips[2]["x"], tips[2]["y"] = nx / 2 - 4.4, ny / 2 + 6.6
    cnt = torch.tensor([3], dtype=torch.int32, device="cuda")
    vec = torch.as_tensor(np.frombuffer(tips.tobytes(), dtype=np.uint8).copy()).cuda()
    want = oracle.sr_integrals(p, u, v, vtu, vtv, wax, way, tips=tips, count=4)
    got = host.sr_integrals(p, du, dv, dvtu, dvtv, ax, ay, tip_count=cnt, tip_vector=vec, count=4)
    assert np.array_equal(got, want)
    cnt0 = torch.tensor([0], dtype=torch.int32, device="cuda")   # empty list keeps tipx0 (defect B3)
    got = host.sr_integrals(p, du, dv, dvtu, dvtv, ax, ay, tip_count=cnt0, tip_vector=vec, count=4)
    assert np.array_equal(got, oracle.sr_integrals(p, u, v, vtu, vtv, wax, way))


@pytest.mark.parametrize("neu,so", [(1, 0), (1, 1), (0, 0), (0, 1)])
@pytest.mark.parametrize("nx,ny", [(96, 64), (130, 101)])
def test_bfecc_bitwise(oracle, neu, so, nx, ny):
    p = oracle.params_default(nx, ny, reduce_sym=True, neumannBC=neu, solidSwitch=so, boundaryVal=0.05)
    u, v = fields(nx, ny, 4)
    solid = (np.random.default_rng(9).uniform(size=(ny, nx)) > 0.15).astype(np.uint8) if so else None
    c, phi = [0.9, -1.3, 0.5], [0.0, 0.0, 0.6]
    wax, way = oracle.cxy_field(p, c, phi, solid=solid)
    want = oracle.advect_bfecc(p, u, v, wax, way, solid=solid)
    du, dv = dev(u), dev(v)
    ds = dev(solid, torch.uint8) if so else None
    uo, vo = torch.empty_like(du), torch.empty_like(dv)
    host.advect_bfecc(p, du, dv, uo, vo, dev(wax), dev(way), solid=ds)
    assert np.array_equal(uo.cpu().numpy(), want[0]) and np.array_equal(vo.cpu().numpy(), want[1])
    # Cxy fused into the advection: same bits, adv field written as a by-product
    ax, ay = torch.zeros(nx * ny, dtype=torch.float64, device="cuda"), torch.zeros(nx * ny, dtype=torch.float64, device="cuda")
    uo2, vo2 = torch.empty_like(du), torch.empty_like(dv)
    host.advect_bfecc_cphi(p, du, dv, uo2, vo2, c, phi, adv_x=ax, adv_y=ay, solid=ds)
    assert torch.equal(uo, uo2) and torch.equal(vo, vo2)
    assert np.array_equal(ax.cpu().numpy(), wax) and np.array_equal(ay.cpu().numpy(), way)


@pytest.mark.skipif(not oracle_lib.have_reference(), reason="oracle/_ref not built")
def test_sr_pieces_vs_reference_kernels(oracle):
    ref = oracle_lib.Reference(nofma=True)
    nx = ny = 128
    p = oracle.params_default(nx, ny, reduce_sym=True, tipOffsetX=40, tipOffsetY=40, tipx0=70.0, tipy0=60.0)
    ref.init(p)
    u, v = fields(nx, ny, 5)
    vtu, vtv = fields(nx, ny, 6)
    c, phi = [0.13, -0.21, 0.04], [0.3, -0.1, 0.77]
    rax, ray = ref.cxy(c, phi)
    ax, ay = torch.zeros(nx * ny, dtype=torch.float64, device="cuda"), torch.zeros(nx * ny, dtype=torch.float64, device="cuda")
    host.cxy_field(p, ax, ay, c, phi)
    # device cos/sin (libdevice) vs host libm: <= 1 ulp of the O(1) terms
    assert np.abs(ax.cpu().numpy() - rax).max() < 1e-15 and np.abs(ay.cpu().numpy() - ray).max() < 1e-15
    rint, rsl = ref.slice_trapz(u, v, rax, ray, vtu, vtv, 0.0, 0.0, 0, want_slices=True)
    s = [torch.zeros(nx * ny, dtype=torch.float64, device="cuda") for _ in range(6)]
    s0 = [torch.zeros(nx * ny, dtype=torch.float64, device="cuda") for _ in range(6)]
    host.slice_fields(p, dev(u), dev(v), s, s0, dev(rax), dev(ray))
    for q in range(6):   # slice_kernel is race-free: bitwise
        assert np.array_equal(s[q].cpu().numpy(), rsl[q]) and np.array_equal(s0[q].cpu().numpy(), rsl[6 + q])
    got = host.sr_integrals(p, dev(u), dev(v), dev(vtu), dev(vtv), dev(rax), dev(ray))
    # reference: 256 blocks atomicAdd(double) in arbitrary order -> 1e-12 relative
    assert np.allclose(got, rint, rtol=1e-12, atol=1e-14)
    assert np.array_equal(host.solve_matrix([0, 0, 0], phi, rint), ref.solve_matrix([0, 0, 0], phi, rint))
    # BFECC: the reference kernel races on uf/ub/ue (advFDBFECC.cu:131-144): a single call reads
    # stale scratch across warps.  Its uf is a pure function of g_in, so repeating the call on
    # the same buffers converges (call 2: correct ub/ue, call 3: correct neighbours of ue) to
    # the synchronous three-sweep result -- which must then equal ours BITWISE.
    ru, rv = ref.bfecc(u, v, rax, ray, repeats=3)
    uo, vo = torch.empty(ny, nx, dtype=torch.float64, device="cuda"), torch.empty(ny, nx, dtype=torch.float64, device="cuda")
    host.advect_bfecc(p, dev(u), dev(v), uo, vo, dev(rax), dev(ray))
    assert np.array_equal(uo.cpu().numpy(), ru) and np.array_equal(vo.cpu().numpy(), rv)
    r1u, _ = ref.bfecc(u, v, rax, ray, repeats=1)
    print("BFECC single racy reference call vs synchronous: max |diff| =", np.abs(r1u - ru).max())


def test_sim_driver_trace_batch_and_pacing(oracle, yh):
    nx = ny = 128
    p = oracle.params_default(nx, ny, timeIntOrder=1, lap4=0)
    u0, v0 = synth.cross_field_ic(nx, ny)
    sim = yh.Sim(p, n_sims=1)
    sim.cross_field_ic()
    tr = sim.run(50, trace=True)
    u, v = sim.get_state()
    wu, wv = oracle.rd_advance(p, 50, u0, v0)
    assert np.array_equal(u[0], wu) and np.array_equal(v[0], wv)
    # electrode trace with the one-step lag of main.cu:1040: sample k = state k at param.point
    uu, vv = u0, v0
    for k in range(3):
        assert tr[k, 0, 0] == uu[ny // 2, nx // 2] and tr[k, 0, 1] == vv[ny // 2, nx // 2]
        uu, vv = oracle.rd_step(p, uu, vv)
    t = sim.tips()
    prev = oracle.rd_advance(p, 49, u0, v0)
    want = oracle.tip_track(p, prev[0], wu, t=p.dt * 50)
    assert t.tobytes() == want.tobytes()
    sim.close()
    # batched sweep (C5 protocol, scaled down): per-sheet pacing period, quiescent IC
    periods = np.array([40, 25, 0], dtype=np.int32)
    sim = yh.Sim(p, n_sims=3)
    sim.set_state(np.zeros((3, ny, nx)), np.zeros((3, ny, nx)))
    sim.set_pacing(periods, 6)
    sim.run(64, tb_steps=4)
    u, v = sim.get_state()
    for z, per in enumerate(periods):
        uu, vv = np.zeros((ny, nx)), np.zeros((ny, nx))
        for s in range(64):
            on = per > 0 and (s % per) <= 6
            uu, vv = oracle.rd_step(p, uu, vv, stim_mouse=bool(on))
        assert np.array_equal(u[z], uu) and np.array_equal(v[z], vv), z
    assert u[0].max() > 0.5 and u[2].max() == 0.0
    sim.close()


@pytest.mark.parametrize("kw", [dict(timeIntOrder=1, lap4=0), dict()])
def test_traced_loop_graph_replay(oracle, yh, kw, monkeypatch):
    """The reference's own loop {step; swap; probe every step} (main.cu:879-885, 1040): on small
    sheets the driver replays 64 {probe, step} pairs from a CUDA graph with the sample slot kept on
    the device.  Trace and state are identical to plain launches and to the oracle."""
    nx = ny = 96
    p = oracle.params_default(nx, ny, **kw)
    u0, v0 = synth.cross_field_ic(nx, ny)
    n = 64 * 4 + 37
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("YH_GRAPHS", mode)
        sim = yh.Sim(p, n_sims=2)
        sim.set_state(np.stack([u0, v0]), np.stack([v0, u0]))
        sim.set_point(nx // 8, ny // 2)
        tr = sim.run(n, trace=True)
        out[mode] = (tr, sim.get_state(), sim.count)
        sim.close()
    assert np.array_equal(out["0"][0], out["1"][0]) and out["0"][2] == out["1"][2] == n
    assert np.array_equal(out["0"][1][0], out["1"][1][0]) and np.array_equal(out["0"][1][1], out["1"][1][1])
    uu, vv = u0, v0
    for k in range(3):   # sample k = state k at the electrode (one-step lag of main.cu:1040)
        assert out["1"][0][k, 0, 0] == uu[ny // 2, nx // 8] and out["1"][0][k, 0, 1] == vv[ny // 2, nx // 8]
        uu, vv = oracle.rd_step(p, uu, vv)
    wu, wv = oracle.rd_advance(p, n, u0, v0)
    assert np.array_equal(out["1"][1][0][0], wu) and np.array_equal(out["1"][1][1][0], wv)
    prev = oracle.rd_advance(p, n - 1, u0, v0)
    assert out["1"][0][n - 1, 0, 0] == prev[0][ny // 2, nx // 8]


def test_reference_launch_api_through_the_shim(oracle):
    """The display()-style loop of tests/shim_driver.cu calls reactionDiffusion_wrapper / swapSoA /
    singleCell_wrapper / tip_wrapper with the reference's signatures (hostPrototypes.h:22-57),
    linked against libyolohtli_shim.so: fields, electrode trace, velTan and tips == oracle."""
    import ctypes as C
    import os
    from yolohtli_b200 import _lib
    so = os.path.join(os.path.dirname(_lib.LIB_PATH), "libyh_shimtest.so")
    drv = C.CDLL(so)
    for kw in (dict(timeIntOrder=1, lap4=0), dict()):   # Euler/5-pt and the default RK4+lap4
        nx = ny = 128
        p = oracle.params_default(nx, ny, **kw)
        u0, v0 = synth.cross_field_ic(nx, ny)
        u, v = u0.copy(), v0.copy()
        nsteps = 30
        trace = np.zeros((nsteps, 2))
        vtu = np.zeros(nx * ny)
        tips = np.zeros(4096, dtype=TIP_DTYPE)
        nt = C.c_int(0)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = drv.yh_shimtest_run(C.byref(p), vp(u), vp(v), nsteps, vp(trace), vp(vtu), vp(tips), C.byref(nt))
        assert rc == 0
        wu, wv = oracle.rd_advance(p, nsteps, u0, v0)
        assert np.array_equal(u, wu) and np.array_equal(v, wv)
        pu, pv = oracle.rd_advance(p, nsteps - 1, u0, v0)
        # singleCell_wrapper reads gateOut AFTER the swap = the previous state (main.cu:1040)
        assert trace[-1, 0] == pu[ny // 2, nx // 2] and trace[0, 0] == u0[ny // 2, nx // 2]
        _, _, wvtu, _ = oracle.rd_step(p, pu, pv, velTan=True)
        assert np.array_equal(vtu, wvtu.ravel())
        want = oracle.tip_track(p, pu, wu, t=p.dt * nsteps)
        assert nt.value == len(want) and tips[:nt.value].tobytes() == want.tobytes()


@pytest.mark.parametrize("args", ["256 60 default", "512 40 euler", "200 30 default"])
def test_shim_zero_source_change_link(args):
    """tests/shim_main.cu: a program that defines the reference's global `paramVar param` like main.cu:37 and
    never calls yh_shim_configure -- the wrappers read `param` (weak reference) -- bitwise == the C ABI."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "yolohtli_b200", "lib", "yh_shim_main")
    assert os.path.exists(exe), "build the library first"
    r = subprocess.run([exe] + args.split(), capture_output=True, text=True, timeout=300)
    print(r.stdout.strip())
    assert r.returncode == 0 and "shim_main PASS" in r.stdout, r.stdout + r.stderr


def test_symmetry_reduction_step_loop(oracle, yh):
    """C3 (scaled down): the display() symmetry-reduction loop (main.cu:894-954) in the headless
    driver == the same loop composed from oracle pieces, bit for bit: fields, (c, phi) history."""
    nx = ny = 128
    p = oracle.params_default(nx, ny, reduce_sym=True, tipOffsetX=40, tipOffsetY=40, tipx0=60.0, tipy0=66.0)
    u0 = wavy(nx, ny) - 0.05
    v0 = 0.3 * wavy(nx, ny, 0.21, 0.17, 1.3)
    nsteps = 14
    sim = yh.Sim(p)
    sim.set_state(u0[None], v0[None])
    rec = sim.run_sr(nsteps)
    gu, gv = sim.get_state()
    gc, gphi = sim.sr_state()
    sim.close()
    u, v = u0.copy(), v0.copy()
    c, phi = np.zeros(3), np.zeros(3)
    ax, ay = np.zeros(nx * ny), np.zeros(nx * ny)
    tip_steps = 0
    for count in range(nsteps):
        us, vs, vtu, vtv = oracle.rd_step(p, u, v, velTan=True)
        tips = oracle.tip_track(p, us, u, t=p.dt * count)
        tip_steps += len(tips) > 0
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
    assert tip_steps > 0, "the test fields should produce tips so the tip-centred disc is exercised"
    assert np.array_equal(gu[0], u) and np.array_equal(gv[0], v)
    assert np.array_equal(gc, c) and np.array_equal(gphi, phi)


def test_symmetry_reduction_device_resident_solve(oracle, yh):
    """yh_sim_run_sr_device keeps the integrals, the 3x3 solve and the frame update on the GPU (no
    host sync per step) and replays chunks of 8 steps from a CUDA graph (step counter, tip time tag
    and record slot live on the device).  Only cos/sin(phi.t) differ (libdevice vs libm, <= 1-2
    ulp): the drift history, the frame and the fields follow the host-solve path to 1e-10 over 100
    steps, the record has the same layout, and runs can be split and mixed with the host path."""
    nx = ny = 128
    p = oracle.params_default(nx, ny, reduce_sym=True, tipOffsetX=40, tipOffsetY=40, tipx0=60.0, tipy0=66.0)
    u0 = wavy(nx, ny) - 0.05
    v0 = 0.3 * wavy(nx, ny, 0.21, 0.17, 1.3)
    nsteps = 100
    a = yh.Sim(p)
    a.set_state(u0[None], v0[None])
    ra = a.run_sr(nsteps)
    au, av = a.get_state()
    ac, aphi = a.sr_state()
    a.close()
    b = yh.Sim(p)
    b.set_state(u0[None], v0[None])
    rb = np.concatenate([b.run_sr_device(60), b.run_sr(5), b.run_sr_device(35)])   # count == 0 inside the first call; graphs in both
    bu, bv = b.get_state()
    bc, bphi = b.sr_state()
    assert b.count == nsteps
    b.close()
    assert np.array_equal(rb[0], ra[0]) and np.array_equal(rb[1], ra[1])   # step 0 runs on the host path
    scale = np.abs(ra).max(axis=0) + 1e-300
    assert (np.abs(rb - ra) / scale).max() < 1e-10
    assert np.abs(bu - au).max() < 1e-10 and np.abs(bv - av).max() < 1e-10
    assert np.allclose(bc, ac, rtol=1e-10, atol=1e-13) and np.allclose(bphi, aphi, rtol=1e-10, atol=1e-13)
    assert np.abs(ra[:, :3]).max() > 1e-3, "the drift must be non-trivial for this comparison to mean anything"


def test_apd_loop_batched_matches_oracle_and_reference(oracle, yh):
    """C5 protocol (scaled down): paced sheets, sAPD every step with the reference's argument order
    (main.cu:1035).  Batched driver == oracle composition == the reference's own loop, bitwise."""
    nx = ny = 96
    p = oracle.params_default(nx, ny, timeIntOrder=1, lap4=0, eps=0.2, beta=3.0)   # short action potentials
    area = synth.stim_area_square(nx, ny)
    periods = np.array([400, 260], dtype=np.int32)
    nsteps, dur = 700, 10
    sim = yh.Sim(p, n_sims=2)
    sim.set_state(np.zeros((2, ny, nx)), np.zeros((2, ny, nx)))
    sim.set_pacing(periods, dur)
    sim.run_apd(nsteps, stim_area=area)
    gu, gv = sim.get_state()
    a1, a2 = sim.get_apd()
    sim.close()
    for z, per in enumerate(periods):
        u, v = np.zeros((ny, nx)), np.zeros((ny, nx))
        st = {k: np.zeros(nx * ny) for k in ("APD1", "APD2", "sAPD", "dAPD", "back", "front")}
        first = np.zeros(nx * ny, dtype=np.uint8)
        import ctypes as C
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        for s in range(nsteps):
            un, vn = oracle.rd_step(p, u, v, stim_mouse=bool((s % per) <= dur))
            # (uold, unew) := (gateIn = new, gateOut = old) after the swap, count = s + 1
            rc = oracle.l.yho_sapd(C.byref(p), s + 1, vp(un), vp(u), vp(st["APD1"]), vp(st["APD2"]), vp(st["sAPD"]),
                                   vp(st["dAPD"]), vp(st["back"]), vp(st["front"]), vp(first), vp(area.reshape(-1)), 1)
            assert rc == 0
            u, v = un, vn
        assert np.array_equal(gu[z], u) and np.array_equal(gv[z], v), z
        assert np.array_equal(a1[z].ravel(), st["APD1"]) and np.array_equal(a2[z].ravel(), st["APD2"]), z
        if oracle_lib.have_reference():
            ref = oracle_lib.Reference(nofma=True)
            ref.init(p)
            ru, rv, r1, r2, _ = ref.apd_run(np.zeros((ny, nx)), np.zeros((ny, nx)), nsteps, int(per), dur, area)
            assert np.array_equal(ru, u) and np.array_equal(r1, st["APD1"]) and np.array_equal(r2, st["APD2"]), z
    assert np.abs(a1).max() > 0, "an action potential should have completed so APD1 is exercised"


# ---- contours and frame colouring ----------------------------------------
def contour_fields(nx, ny, mode, seed=3):
    """Inputs with plenty of crossings; the LAST ROW holds none (there the reference reads past
    the end of the array, spaceAPD.cu:52-53)."""
    X, Y = np.meshgrid(np.arange(nx, dtype=float), np.arange(ny, dtype=float))
    rng = np.random.default_rng(seed)
    area = (rng.uniform(size=(ny, nx)) > 0.2).astype(np.uint8)
    if mode == 1:   # sAPD: sign field with a few exact zeros (spaceAPD.cu:343)
        s = np.sign(np.sin(0.23 * X + 0.4) * np.cos(0.19 * Y) + 0.3 * np.sin(0.05 * X * Y / nx))
        s[rng.uniform(size=(ny, nx)) < 0.01] = 0.0
        s[-1, :] = 1.0
        return None, s, area
    u = 0.55 + 0.5 * np.sin(0.13 * X) * np.cos(0.17 * Y + 0.3)
    v = 0.9 + 0.85 * np.sin(0.21 * X + 0.1 * Y) + 0.02 * rng.normal(size=(ny, nx))
    u[-1, :] = 1.0
    return u, v, area


def gpu_contour(p, f1, f2, mode, t=0.0, area=None, capacity=None):
    n = p.nx * p.ny
    cap = capacity if capacity is not None else 2 * n
    cnt = torch.full((1,), -1, dtype=torch.int32, device="cuda")
    vec = torch.zeros(cap * 12, dtype=torch.uint8, device="cuda")
    pl = torch.full((n,), 7, dtype=torch.uint8, device="cuda")   # must be reset by the call
    host.contour(p, dev(f1) if f1 is not None else None, dev(f2), cnt, vec, mode, contour_plot=pl,
                 stimArea=dev(area, torch.uint8) if area is not None else None, t=t, capacity=cap)
    torch.cuda.synchronize()
    pts, count = host.contour_to_numpy(cnt, vec)
    return pts, count, pl.cpu().numpy()


@pytest.mark.parametrize("nx,ny", [(48, 40), (333, 257), (1024, 1024)])
@pytest.mark.parametrize("mode", [1, 2, 3])
def test_contour_vs_oracle_bitwise(oracle, nx, ny, mode):
    p = oracle.params_default(nx, ny)
    f1, f2, area = contour_fields(nx, ny, mode)
    for a in (area, None):
        want, wn, wplot = oracle.contour(p, f1, f2, mode, t=12.5, stimArea=a, plot=True)
        got, gn, gplot = gpu_contour(p, f1, f2, mode, t=12.5, area=a)
        assert wn > nx // 4 and gn == wn
        assert got.tobytes() == want.tobytes()            # same points, same (canonical) order
        assert np.array_equal(gplot, wplot)
    # a second call on the same buffers (ticket re-armed, epoch bumped), and a clipped list
    got2, gn2, _ = gpu_contour(p, f1, f2, mode, t=12.5, area=None, capacity=5)
    assert gn2 == wn and got2.tobytes() == want[:5].tobytes()


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_contour_vs_reference_kernels(oracle, mode):
    """The reference appends by atomicAdd (order = scheduling): compare as sorted multisets; the
    raster must be identical."""
    if not oracle_lib.have_reference():
        pytest.skip("oracle/_ref not built")
    nx, ny = 333, 257
    p = oracle.params_default(nx, ny)
    ref = oracle_lib.Reference(nofma=True)
    ref.init(p)
    f1, f2, area = contour_fields(nx, ny, mode)
    rpts, rplot = ref.contour(f1, f2, mode, t=3.5, stimArea=area)
    got, gn, gplot = gpu_contour(p, f1, f2, mode, t=3.5, area=area)
    assert gn == len(rpts) > 50
    assert np.array_equal(np.sort(got, order=("y", "x", "t")), np.sort(rpts, order=("y", "x", "t")))
    assert np.array_equal(gplot, rplot)


def test_contour_golden_fixtures(oracle):
    import glob
    import os
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "contour_*.npz")))
    if not files:
        pytest.skip("no contour fixtures yet")
    for f in files:
        g = np.load(f)
        ny, nx = g["f2"].shape
        p = oracle.params_default(nx, ny)
        f1 = g["f1"] if int(g["mode"]) != 1 else None
        got, gn, gplot = gpu_contour(p, f1, g["f2"], int(g["mode"]), t=float(g["t"]), area=g["area"])
        ref = g["pts"].view(oracle_lib.CONTOUR_DTYPE).reshape(-1)
        assert np.array_equal(np.sort(got, order=("y", "x", "t")), np.sort(ref, order=("y", "x", "t")))
        assert np.array_equal(gplot, g["plot"].reshape(-1))


def test_rgba_vs_oracle(oracle, yh):
    nx, ny = 300, 200
    p = oracle.params_default(nx, ny)
    rng = np.random.default_rng(9)
    field = rng.uniform(-0.4, 1.4, (ny, nx))                 # leaves the colour range on both sides
    lines = (rng.uniform(size=(ny, nx)) < 0.05).astype(np.uint8)
    cmap = yh.io.cmap_read(None, capacity=500)
    want = oracle.rgba(p, field, cmap, -0.1, 1.1, lines=lines)
    out = torch.zeros(nx * ny, dtype=torch.int32, device="cuda")
    host.rgba(p, dev(field), out, dev(cmap.view(np.int32), torch.int32), -0.1, 1.1, lines=dev(lines, torch.uint8))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want)
    host.rgba(p, dev(field), out, dev(cmap.view(np.int32), torch.int32), -0.1, 1.1)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), oracle.rgba(p, field, cmap, -0.1, 1.1))
