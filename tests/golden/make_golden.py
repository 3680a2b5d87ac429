"""Generates (on a GPU box, from the reference's OWN kernels in oracle/_ref built with
--fmad=false) and checks (anywhere, against the plain-C oracle) the golden fixtures of the
hot path.  The reference ships no golden vectors of its own.

    python -m tests.golden.make_golden          # on the B200 box; writes tests/golden/*.npz

Fixtures are small (<= 48 x 40 cells) so they can be committed.  Only race-free modes of the
reference are recorded as bitwise fixtures (Euler, lap4 = 0, every boundary/mask branch);
tips, sAPD, Cxy, slice/trapz and solve_matrix are deterministic in the reference and are
recorded too (tip lists sorted, integrals to 1e-12 because of atomicAdd ordering)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PARAM_KEYS = ("solidSwitch", "neumannBC", "gateDiff", "anisotropy", "lap4", "timeIntOrder")
NX, NY = 48, 40


def _params(oracle, g_or_kw):
    kw = {k: int(g_or_kw[k]) for k in PARAM_KEYS if k in g_or_kw}
    rs = int(g_or_kw["reduce_sym"]) if "reduce_sym" in g_or_kw else 0
    p = oracle.params_default(NX, NY, reduce_sym=bool(rs), **kw)
    for k in ("tipOffsetX", "tipOffsetY"):
        if k in g_or_kw:
            setattr(p, k, int(g_or_kw[k]))
    for k in ("tipx0", "tipy0"):
        if k in g_or_kw:
            setattr(p, k, float(g_or_kw[k]))
    return p


def _sorted_tips(t):
    key = np.floor(t["x"]).astype(np.int64) + NX * np.floor(t["y"]).astype(np.int64)
    return t[np.argsort(key, kind="stable")]


def check_oracle(oracle, g):
    kind = str(g["kind"])
    p = _params(oracle, g)
    if kind == "rd":
        solid = g["solid"] if int(g["solidSwitch"]) else None
        u, v = oracle.rd_advance(p, int(g["nsteps"]), g["u0"], g["v0"], solid=solid,
                                 stim_mouse=bool(g["stim"]), point=(int(g["px"]), int(g["py"])))
        assert np.array_equal(u, g["u"]) and np.array_equal(v, g["v"]), "rd fields differ bitwise"
        uo, vo, vtu, vtv = oracle.rd_step(p, g["u0"], g["v0"], solid=solid, velTan=True,
                                          stim_mouse=bool(g["stim"]), point=(int(g["px"]), int(g["py"])))
        if int(g["gateDiff"]):
            assert np.array_equal(vtu.reshape(-1), g["vtu1"].reshape(-1))
    elif kind == "tip":
        t = oracle.tip_track(p, g["past"], g["present"], t=float(g["t"]), algorithm=int(g["algorithm"]))
        ref = _sorted_tips(g["tips"].view(t.dtype).reshape(-1))
        assert len(t) == len(ref)
        # '+'/'-' roots of one cell may swap under the sort; compare as multisets per cell
        a = np.sort(np.stack([t["x"], t["y"]], 1).view("f4,f4").reshape(-1))
        b = np.sort(np.stack([ref["x"], ref["y"]], 1).view("f4,f4").reshape(-1))
        assert np.array_equal(a, b)
    elif kind == "sapd":
        st = oracle.sapd_sequence(p, list(g["seq"]), count0=int(g["count0"]), stimArea=g["area"],
                                  stimulate=bool(g["stimulate"]))
        for k in ("APD1", "APD2", "sAPD", "back", "front"):
            assert np.array_equal(st[k], g[k]), k
        assert np.array_equal(st["first"], g["first"])
    elif kind == "sr":
        c, phi = g["c"], g["phi"]
        ax, ay = oracle.cxy_field(p, c, phi)
        # libdevice cos/sin vs glibc: last-ulp differences allowed
        assert np.allclose(ax, g["ax"], rtol=0, atol=1e-15) and np.allclose(ay, g["ay"], rtol=0, atol=1e-15)
        s, s0 = oracle.slice(p, g["u"], g["v"], g["ax"], g["ay"])
        ref = g["slices"]
        for q in range(6):
            assert np.array_equal(s[q], ref[q]), f"slice {q}"
            assert np.array_equal(s0[q], ref[6 + q]), f"slice0 {q}"
        I = oracle.trapz(p, s, s0, g["vtu"], g["vtv"])
        assert np.allclose(I, g["integrals"], rtol=1e-12, atol=1e-15)
        assert np.array_equal(oracle.solve_matrix([0, 0, 0], phi, g["integrals"]), g["csol"])
    elif kind == "bfecc":
        uo, vo = oracle.advect_bfecc(p, g["u"], g["v"], g["ax"], g["ay"])
        # reference kernel called 3x on the same buffers = its synchronous fixed point (see ref_harness.cu)
        assert np.array_equal(uo, g["uo"]) and np.array_equal(vo, g["vo"])
    elif kind == "contour":
        from tests import oracle_lib
        ny, nx = g["f2"].shape
        p = oracle.params_default(nx, ny)
        mode = int(g["mode"])
        pts, n, plot = oracle.contour(p, g["f1"] if mode != 1 else None, g["f2"], mode, t=float(g["t"]),
                                      stimArea=g["area"], plot=True)
        ref = g["pts"].view(oracle_lib.CONTOUR_DTYPE).reshape(-1)
        key = ("y", "x", "t")   # the reference appends by atomicAdd: compare as sorted multisets
        assert n == len(ref) and np.array_equal(np.sort(pts, order=key), np.sort(ref, order=key))
        assert np.array_equal(plot, g["plot"].reshape(-1))
    else:
        raise AssertionError(kind)


def contour_fixtures():
    """python -m tests.golden.make_golden contour  (on the B200 box)."""
    from tests import oracle_lib
    from tests.test_gpu_aux import contour_fields
    oracle = oracle_lib.load()
    ref = oracle_lib.Reference(nofma=True)
    ref.init(oracle.params_default(NX, NY))
    for mode in (1, 2, 3):
        f1, f2, area = contour_fields(NX, NY, mode, seed=40 + mode)
        pts, plot = ref.contour(f1, f2, mode, t=7.25, stimArea=area)
        name = os.path.join(HERE, f"contour_{mode}.npz")
        np.savez_compressed(name, kind="contour", mode=mode, t=7.25, f1=f1 if f1 is not None else np.zeros(1),
                            f2=f2, area=area, pts=pts.view(np.uint8), plot=plot)
        check_oracle(oracle, np.load(name))
        print("  oracle == reference:", os.path.basename(name), len(pts), "points")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "contour":
        return contour_fixtures()
    from tests import oracle_lib
    from yolohtli_b200 import synth
    oracle = oracle_lib.load()
    ref = oracle_lib.Reference(nofma=True)
    rng = np.random.default_rng(20261017)
    u0 = rng.uniform(-0.1, 1.1, (NY, NX))
    v0 = rng.uniform(0.0, 1.0, (NY, NX))
    solid = (rng.uniform(size=(NY, NX)) > 0.15).astype(np.uint8)
    n = 0
    for neu in (1, 0):
        for so in (0, 1):
            for gd in (1, 0):
                for an in (0, 1):
                    if neu and so and an:
                        continue   # anisotropy is ignored in that branch (reactionDiffusion.cu:184)
                    if (not neu) and so:
                        continue   # Dirichlet+solid indexes out of bounds at the edges (UB)
                    kw = dict(solidSwitch=so, neumannBC=neu, gateDiff=gd, anisotropy=an, lap4=0, timeIntOrder=1)
                    p = oracle.params_default(NX, NY, **kw)
                    if an:
                        oracle.l.yho_params_derive(__import__("ctypes").byref(p), __import__("ctypes").c_double(0.001),
                                                   __import__("ctypes").c_double(0.0004), __import__("ctypes").c_double(0.0002))
                    ref.init(p)
                    stim = int(n % 2 == 0)
                    u, v, _ = ref.rd_run(u0, v0, 5, solid=solid, stim_mouse=bool(stim), point=(20, 18))
                    _, _, vtu1, vtv1, _ = ref.rd_run(u0, v0, 1, solid=solid, stim_mouse=bool(stim), point=(20, 18), velTan=True)
                    if an:
                        continue   # derived constants differ from params_default; covered by GPU tests
                    np.savez_compressed(os.path.join(HERE, f"rd_{n:02d}.npz"), kind="rd", nsteps=5, u0=u0, v0=v0,
                                        solid=solid, stim=stim, px=20, py=18, u=u, v=v, vtu1=vtu1, vtv1=vtv1, **kw)
                    n += 1
    # tips: a real spiral-ish pair of fields
    p = oracle.params_default(NX, NY, timeIntOrder=1, lap4=0)
    X, Y = np.meshgrid(np.arange(NX, dtype=float), np.arange(NY, dtype=float))
    present = 0.7 + 0.3 * np.sin(0.31 * X + 0.1) * np.cos(0.27 * Y)
    past = 0.7 + 0.3 * np.cos(0.29 * X) * np.sin(0.33 * Y + 0.2)
    for alg in (1, 2):
        ref.init(p)
        tips = ref.tip(present, past, t=3.25, algorithm=alg)
        np.savez_compressed(os.path.join(HERE, f"tip_alg{alg}.npz"), kind="tip", present=present, past=past,
                            t=3.25, algorithm=alg, tips=tips.view(np.uint8), lap4=0, timeIntOrder=1)
    # sAPD
    p = oracle.params_default(NX, NY, timeIntOrder=1, lap4=0)
    ref.init(p)
    seq = [np.clip(0.5 + 0.6 * np.sin(0.2 * k + 0.05 * X + 0.07 * Y), -0.1, 1.1) for k in range(70)]
    area = synth.stim_area_square(NX, NY)
    for stimulate in (0, 1):
        st = ref.sapd_sequence(seq, count0=1, stimArea=area, stimulate=bool(stimulate))
        np.savez_compressed(os.path.join(HERE, f"sapd_{stimulate}.npz"), kind="sapd", seq=np.stack(seq), count0=1,
                            area=area, stimulate=stimulate, lap4=0, timeIntOrder=1, **st)
    # symmetry-reduction pieces
    kw = dict(reduce_sym=1, tipOffsetX=14, tipOffsetY=14, tipx0=25.0, tipy0=19.0)
    p = _params(oracle, kw)
    ref.init(p)
    c, phi = np.array([0.13, -0.21, 0.04]), np.array([0.3, -0.1, 0.77])
    ax, ay = ref.cxy(c, phi)
    vtu, vtv = rng.normal(size=(NY, NX)), rng.normal(size=(NY, NX))
    integrals, slices = ref.slice_trapz(u0, v0, ax, ay, vtu, vtv, 0.0, 0.0, 0, want_slices=True)
    csol = ref.solve_matrix([0, 0, 0], phi, integrals)
    np.savez_compressed(os.path.join(HERE, "sr_00.npz"), kind="sr", u=u0, v=v0, ax=ax, ay=ay, vtu=vtu, vtv=vtv,
                        c=c, phi=phi, integrals=integrals, slices=slices, csol=csol, **kw)
    uo, vo = ref.bfecc(u0, v0, ax, ay, repeats=3)
    np.savez_compressed(os.path.join(HERE, "bfecc_00.npz"), kind="bfecc", u=u0, v=v0, ax=ax, ay=ay, uo=uo, vo=vo, **kw)
    print("golden fixtures written:", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
    # self-check
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            check_oracle(oracle, np.load(os.path.join(HERE, f)))
            print("  oracle == reference:", f)


if __name__ == "__main__":
    main()
