"""CPU: the data formats either side of the hot path (include/yolohtli_io.h):
text written character for character as the reference's printf calls write it, the parse rules
of its readers, and the lossless snapshot."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def c_f(x):
    """C's "%f" of a value that went through a float cast."""
    return "%f" % float(np.float32(x))


def test_io_header_symbols_exported(yh):
    from yolohtli_b200 import _lib
    txt = open(os.path.join(ROOT, "include", "yolohtli_io.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    syms = sorted(set(re.findall(r"\b(yh_io_[a-z0-9_]+)\s*\(", txt)))
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(yh.lib(), s), s
        assert s in _lib.SIGNATURES, s


def test_params_csv_text_and_roundtrip(yh, tmp_path):
    rp = yh.io.run_params_default(512, 512)
    f = tmp_path / "dataparamcsv.csv"
    yh.io.params_write_csv(f, rp)
    lines = open(f).read().split("\n")
    assert lines[-1] == "" and len(lines) - 1 == 91   # printFunctions.cu:293-397: 2 paths + 89 values
    # spot-check against the reference's format strings (printFunctions.cu:293-397)
    assert lines[0] == "Initial condition path:,NA" and lines[1] == "Results file path:,NA"
    assert lines[3] == "plot tip on screen:,1"
    assert lines[14] == "Tip trajectory algorithm:, 1"          # ", %d" with the space, :307
    assert lines[18] == "Laplacian order:,1"                     # lap4 = 4 printed as a bool, :311
    assert lines[19] == "Scheme order for time integration (Euler or RK4):,4"
    assert lines[22] == "# grid points X =,512"
    assert lines[26] == "Physical dx," + c_f(12.0 / 511.0)
    assert lines[28] == "Time step:,0.020000"
    assert lines[38] == "rx (Dxx*dt/(dx*dx)):," + c_f(0.02 * 0.001 / (12.0 / 511.0) ** 2)
    assert lines[50] == "Electrode position x:,256.000000"
    assert lines[52] == "Stimulus period (ms):,600.000000"
    assert lines[-2] == "theta:,0.000000" and lines[-5] == "delta:,1.000000"
    # reader: positional, every value through strtof (saveFiles.cu:577), derived scalars recomputed
    q = yh.io.params_read_csv(f)
    assert (q.k.nx, q.k.ny, q.k.timeIntOrder, q.k.lap4, q.k.tipAlgorithm) == (512, 512, 4, 1, 1)
    assert q.k.dt == float(np.float32(0.02)) and q.k.hx == float(np.float32("%f" % (12.0 / 511.0)))
    assert q.k.rx == q.k.dt * q.Dxx / (q.k.hx * q.k.hx)
    assert q.contourMode == 3 and q.point_x == 256 and q.k.tipOffsetX == 160
    assert q.contourThresh2 == float(np.float32(0.85)) and q.k.Uth == float(np.float32(0.7))
    # a second write of what was read reproduces the text, except for the derived scalars, which
    # the reader recomputes from the 6-decimal hx / dt (main.cu:148-155 does the same)
    g = tmp_path / "again.csv"
    yh.io.params_write_csv(g, q)
    diff = [a.split(",")[0] for a, b in zip(open(g).read().split("\n"), lines) if a != b]
    assert all(d.split(" ")[0] in ("rx", "ry", "rxy", "rbx", "rby", "invdx", "invdy", "qx4", "qy4", "fx4", "fy4")
               for d in diff), diff
    # symmetry reduction: dt halved in memory (main.cu:148), written doubled (printFunctions.cu:325)
    rp.reduceSym = 1
    rp.k.dt = 0.01
    yh.io.params_write_csv(f, rp)
    assert open(f).read().split("\n")[28] == "Time step:,0.020000"
    q = yh.io.params_read_csv(f)
    assert q.reduceSym == 1 and q.k.dt == 0.5 * float(np.float32(0.02))
    with pytest.raises(yh.YolohtliError):
        yh.io.params_read_csv(tmp_path / "missing.csv")
    open(g, "w").write("a,1\nb,2\n")
    with pytest.raises(yh.YolohtliError):
        yh.io.params_read_csv(g)


def test_state_text_window_and_snapshot(yh, tmp_path):
    rng = np.random.default_rng(5)
    nx, ny = 37, 23
    u, v = rng.uniform(-0.1, 1.1, (ny, nx)), rng.uniform(0, 1, (ny, nx))
    f = tmp_path / "raw_data.dat"
    yh.io.state_write_text(f, u, v)
    lines = open(f).read().split("\n")
    assert len(lines) == nx * ny + 1
    assert lines[0] == c_f(u[0, 0]) + " " + c_f(v[0, 0])           # "%f %f\n", i fastest
    assert lines[nx + 2] == c_f(u[1, 2]) + " " + c_f(v[1, 2])
    ru, rv = yh.io.state_read_text(f, nx, ny)
    want = np.array([float(np.float32("%f" % float(np.float32(x)))) for x in u.ravel()]).reshape(ny, nx)
    assert np.array_equal(ru, want) and np.abs(ru - u).max() < 1e-6 and np.abs(rv - v).max() < 1e-6
    # loadData's own format ("%f\t%f", saveFiles.cu:529) parses the same
    open(f, "w").write("".join("%f\t%f\n" % (float(np.float32(a)), float(np.float32(b))) for a, b in zip(u.ravel(), v.ravel())))
    tu, _ = yh.io.state_read_text(f, nx, ny)
    assert np.array_equal(tu, ru)
    open(f, "w").write("0.5 0.25\n")
    with pytest.raises(yh.YolohtliError):
        yh.io.state_read_text(f, nx, ny)
    # sub-window around a tip: [floor(t)-off-1, floor(t)+off+1), clipped to the sheet
    n = yh.io.state_write_window(f, u, v, 10.7, 9.2, 3, 2)
    assert n == 8 * 6
    got = np.loadtxt(f)
    assert np.allclose(got[:, 0].reshape(6, 8), u[6:12, 6:14], atol=1e-6)
    assert yh.io.state_write_window(f, u, v, 1.0, 1.0, 3, 3) == 5 * 5   # columns -3..-1 clipped
    # lossless snapshot, batched
    U, V = rng.normal(size=(3, ny, nx)), rng.normal(size=(3, ny, nx))
    U[0, 0, 0] = -0.0
    s = tmp_path / "state.yhsnap"
    yh.io.snapshot_write(s, U, V, count=12345, physical_time=246.9)
    a, b, cnt, t = yh.io.snapshot_read(s)
    assert a.tobytes() == U.tobytes() and b.tobytes() == V.tobytes() and cnt == 12345 and t == 246.9
    assert os.path.getsize(s) == 64 + 2 * 8 * U.size
    open(s, "r+b").write(b"NOTASNAP")
    with pytest.raises(yh.YolohtliError):
        yh.io.snapshot_read(s)


def test_mask_parse_rule(yh, tmp_path):
    """main.cu:676-680: fscanf("%f") then value > 0.5 => tissue -- bit-exact, any float syntax."""
    f = tmp_path / "holes4.dat"
    open(f, "w").write("1.000000e+00\n0.000000e+00\n0.5\n0.50001\n1\n0\n-3\n7.5e-1 0.49999 1e0\n2 0.500000029\n")
    m = yh.io.mask_read(f, 4, 3)
    assert m.ravel().tolist() == [1, 0, 0, 1, 1, 0, 0, 1, 0, 1, 1, 0]   # 0.500000029 parses to 0.5f exactly (a double parse would say tissue)
    from yolohtli_b200 import synth
    big = synth.hole_mask(96, seed=11)
    g = tmp_path / "holes96.dat"
    yh.io.mask_write(g, big)
    assert np.array_equal(yh.io.mask_read(g, 96, 96), big)
    assert np.array_equal(synth.read_mask_dat(str(g), 96), big)          # the Python parser agrees
    assert open(g).readline() == "%e\n" % float(big.ravel()[0])
    open(g, "w").write("1 1 1\n")
    with pytest.raises(yh.YolohtliError):
        yh.io.mask_read(g, 96, 96)


def test_domain_objects(yh):
    """intglArea / stimArea / stimulus of domainObjects (main.cu:686-848), float-narrowed coordinates."""
    from yolohtli_b200 import synth
    rp = yh.io.run_params_default(512, 512)
    ia, sa, st = yh.io.domain_objects(rp)
    assert np.array_equal(sa, synth.stim_area_square(512, 512))          # square domain: rows j >= 35
    f32 = np.float32
    i = np.arange(512)
    x0 = (i.astype(f32).astype(np.float64) * rp.k.hx - 0.5 * rp.k.Lx).astype(f32).astype(np.float64)
    X, Y = np.meshgrid(x0, x0)
    rt = 0.5 * ((rp.k.tipOffsetX + rp.k.tipOffsetY) * rp.k.hx)
    assert np.array_equal(ia, ((X.astype(f32) * X.astype(f32) + Y.astype(f32) * Y.astype(f32)).astype(np.float64) < rt * rt).astype(np.uint8))
    inside = ((X - rp.stcx) ** 2 + (Y - rp.stcy) ** 2) < rp.rdomStim ** 2
    assert np.array_equal(st, np.where(inside, rp.stimMag, 0.0)) and 300 < inside.sum() < 900
    rp.k.solidSwitch = 1
    _, sa, _ = yh.io.domain_objects(rp)
    assert np.array_equal(sa, (~(((X - rp.stcx) ** 2 + (Y - rp.stcy) ** 2) < rp.rdomAPD ** 2)).astype(np.uint8))


def test_series_writers(yh, tmp_path):
    from yolohtli_b200.host import CONTOUR_DTYPE, TIP_DTYPE
    tips = np.array([(10.25, 20.5, 0.0, 0.0, 1.5), (300.125, 7.0, -1.0, 2.0, 1.5)], dtype=TIP_DTYPE)
    p1, p2 = tmp_path / "dataTip.dat", tmp_path / "dataTipSize.dat"
    yh.io.tips_append(p1, p2, tips, first=True)
    yh.io.tips_append(p1, p2, tips[:0])               # empty sample: nothing written (printFunctions.cu:181)
    yh.io.tips_append(p1, p2, tips[:1])
    assert open(p1).read() == ("10.250000 20.500000 0.000000 0.000000 1.500000\n"
                               "300.125000 7.000000 -1.000000 2.000000 1.500000\n"
                               "10.250000 20.500000 0.000000 0.000000 1.500000\n")
    assert open(p2).read() == "2\n1\n"
    yh.io.tips_append(p1, p2, tips[:1], first=True)   # first call truncates
    assert open(p2).read() == "1\n"
    pts = np.array([(1.5, 2.0, 0.25), (3.0, 4.5, 0.25)], dtype=CONTOUR_DTYPE)
    yh.io.contour_append(p1, p2, pts, first=True)
    assert open(p1).read() == "1.500000 2.000000 0.250000\n3.000000 4.500000 0.250000\n" and open(p2).read() == "2\n"
    cphi = np.array([[0.1, -0.2, 0.03, 1.0, 2.0, 0.5], [0.11, -0.21, 0.031, 1.1, 2.1, 0.51]])
    s = tmp_path / "c_phi_list_sym.dat"
    yh.io.sym_write(s, cphi)
    assert open(s).read() == ("0.100000 -0.200000 0.030000 1.000000 2.000000 0.500000\n"
                              "0.110000 -0.210000 0.031000 1.100000 2.100000 0.510000\n")
    e = tmp_path / "electrode.dat"
    yh.io.series_write(e, [0.5, 0.75], [0.1, 0.2], 0.02, 50)
    assert open(e).read() == "0.000000\t0.500000\t0.100000\t1.000000\t0.750000\t0.200000\t"
    yh.io.contour_length_write(e, [12, 40], 0.02, 50)
    assert open(e).read() == "0.000000\t12.000000\t1.000000\t40.000000\t"
    # reconstruction of the original-frame tip path (DATA/processSymmetry.m:68-89)
    x, y = np.array([100.5, 101.25], dtype=np.float32), np.array([200.0, 199.5], dtype=np.float32)
    dx = dy = 12.0 / 511.0
    X, Y = yh.io.reconstruct_tip(x, y, cphi, dx, dy)
    xt, yt = (x.astype(np.float64) - 1) * dx, (y.astype(np.float64) - 1) * dy
    ph = cphi[:, 5]
    assert np.allclose(X, -cphi[:, 3] - yt * np.sin(-ph) + xt * np.cos(-ph), rtol=0, atol=1e-15)
    assert np.allclose(Y, -cphi[:, 4] + xt * np.sin(-ph) + yt * np.cos(-ph), rtol=0, atol=1e-15)


def test_colour_map_and_frame(yh, tmp_path):
    f = tmp_path / "cmap.dat"
    open(f, "w").write("\t   3      \n\t0.0000    0.6500    0.6500\n    1.0 0.5 0.25\n 0.2 0.4 1.0\n")
    cm = yh.io.cmap_read(f)
    # main.cu:1461-1464: 255<<24 | (int)(b*255.0f)<<16 | (int)(g*255.0f)<<8 | (int)(r*255.0f)
    f32 = np.float32
    want = [(255 << 24) | (int(f32(b) * f32(255)) << 16) | (int(f32(g) * f32(255)) << 8) | int(f32(r) * f32(255))
            for r, g, b in [(0.0, 0.65, 0.65), (1.0, 0.5, 0.25), (0.2, 0.4, 1.0)]]
    assert cm.tolist() == want
    with pytest.raises(yh.YolohtliError):
        yh.io.cmap_read(f, capacity=2)
    ramp = yh.io.cmap_read(None, capacity=64)
    assert len(ramp) == 64 and (ramp >> 24 == 255).all() and len(set(ramp.tolist())) == 64
    img = np.array([[want[0], want[1]], [want[2], 0]], dtype=np.uint32)
    g = tmp_path / "frame.ppm"
    yh.io.frame_write_ppm(g, img)
    raw = open(g, "rb").read()
    assert raw.startswith(b"P6\n2 2\n255\n") and len(raw) == 11 + 12
    px = np.frombuffer(raw[11:], dtype=np.uint8).reshape(2, 2, 3)
    assert px[1, 0].tolist() == [want[0] & 255, (want[0] >> 8) & 255, (want[0] >> 16) & 255]   # row 0 at the bottom
    assert px[0, 1].tolist() == [0, 0, 0]
