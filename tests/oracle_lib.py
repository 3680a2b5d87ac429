"""ctypes binding of the plain-C oracle (oracle/libyh_oracle.so) and of the reference's own
kernels built headless (oracle/_ref/libyhref*.so).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libyh_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libyhref.so")
REF_NOFMA_SO = os.path.join(ORACLE_DIR, "_ref", "libyhref_nofma.so")

from yolohtli_b200._lib import YhParams, YhTip  # noqa: E402  (struct layouts only)

TIP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("vx", "<f4"), ("vy", "<f4"), ("t", "<f4")])
CONTOUR_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("t", "<f4")])
CONTOUR_THRESH = (0.8, 0.85, 0.7)   # saveFiles.cu:215-217
_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
_P = C.POINTER(YhParams)


def build():
    subprocess.check_call(["make", "-C", ORACLE_DIR, "libyh_oracle.so"], stdout=subprocess.DEVNULL)


def _np(a, dtype=np.float64):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    def __init__(self, path=ORACLE_SO):
        if not os.path.exists(path):
            build()
        self.l = C.CDLL(path)
        self.l.yho_get_threads.restype = _i

    def params_default(self, nx=512, ny=512, reduce_sym=False, scale_L=False, **over):
        p = YhParams()
        rc = self.l.yho_params_default(C.byref(p), nx, ny, int(reduce_sym), int(scale_L))
        assert rc == 0
        for k, v in over.items():
            setattr(p, k, v)
        return p

    def set_threads(self, n):
        self.l.yho_set_threads(int(n))

    def threads(self):
        return self.l.yho_get_threads()

    def rd_step(self, p, u, v, solid=None, stim_mouse=False, point=None, velTan=False):
        u, v = _np(u), _np(v)
        uo, vo = np.empty_like(u), np.empty_like(v)
        vtu = np.zeros_like(u) if velTan else None
        vtv = np.zeros_like(u) if velTan else None
        px, py = point if point is not None else (p.nx // 2, p.ny // 2)
        s = _np(solid, np.uint8) if solid is not None else None
        rc = self.l.yho_rd_step(C.byref(p), _p(u), _p(v), _p(uo), _p(vo), _p(vtu), _p(vtv), _p(s),
                                int(stim_mouse), px, py)
        assert rc == 0, rc
        return (uo, vo, vtu, vtv) if velTan else (uo, vo)

    def rd_advance(self, p, nsteps, u, v, solid=None, stim_mouse=False, point=None):
        u, v = _np(u).copy(), _np(v).copy()
        px, py = point if point is not None else (p.nx // 2, p.ny // 2)
        s = _np(solid, np.uint8) if solid is not None else None
        rc = self.l.yho_rd_advance(C.byref(p), nsteps, _p(u), _p(v), _p(s), int(stim_mouse), px, py)
        assert rc == 0, rc
        return u, v

    def tip_track(self, p, u_past, u_present, t=0.0, algorithm=None, capacity=65536, plot=False):
        a, b = _np(u_past), _np(u_present)
        vec = np.zeros(capacity, dtype=TIP_DTYPE)
        cnt = C.c_int(0)
        pl = np.zeros(p.nx * p.ny, dtype=np.uint8) if plot else None
        alg = algorithm if algorithm is not None else p.tipAlgorithm
        rc = self.l.yho_tip_track(C.byref(p), _p(a), _p(b), _p(pl), C.byref(cnt), _p(vec), capacity,
                                  _d(t), alg)
        assert rc == 0, rc
        out = vec[: cnt.value].copy()
        return (out, pl) if plot else out

    @staticmethod
    def _six(arrs):
        arr = (C.c_void_p * 6)()
        for i, a in enumerate(arrs):
            arr[i] = a.ctypes.data
        return arr

    def slice(self, p, u, v, adv_x, adv_y, scheme=2, start=True, tips=None, count=0):
        u, v, ax, ay = _np(u), _np(v), _np(adv_x), _np(adv_y)
        s = [np.zeros(p.nx * p.ny) for _ in range(6)]
        s0 = [np.zeros(p.nx * p.ny) for _ in range(6)]
        tv = np.ascontiguousarray(tips, dtype=TIP_DTYPE) if tips is not None else None
        n = len(tv) if tv is not None else 0
        rc = self.l.yho_slice(C.byref(p), _p(u), _p(v), self._six(s), self._six(s0), 1, int(start),
                              _p(ax), _p(ay), scheme, n, _p(tv), count)
        assert rc == 0, rc
        return s, s0

    def trapz(self, p, s, s0, vtu, vtv, tips=None, count=0):
        s = [_np(a) for a in s]
        s0 = [_np(a) for a in s0]
        vtu, vtv = _np(vtu), _np(vtv)
        out = np.zeros(12)
        tv = np.ascontiguousarray(tips, dtype=TIP_DTYPE) if tips is not None else None
        n = len(tv) if tv is not None else 0
        rc = self.l.yho_trapz(C.byref(p), self._six(s), self._six(s0), _p(vtu), _p(vtv), _p(out), n,
                              _p(tv), count)
        assert rc == 0, rc
        return out

    def sr_integrals(self, p, u, v, vtu, vtv, adv_x, adv_y, tips=None, count=0):
        u, v, vtu, vtv, ax, ay = map(_np, (u, v, vtu, vtv, adv_x, adv_y))
        out = np.zeros(12)
        tv = np.ascontiguousarray(tips, dtype=TIP_DTYPE) if tips is not None else None
        n = len(tv) if tv is not None else 0
        rc = self.l.yho_sr_integrals(C.byref(p), _p(u), _p(v), _p(vtu), _p(vtv), _p(ax), _p(ay),
                                     _p(out), n, _p(tv), count)
        assert rc == 0, rc
        return out

    def solve_matrix(self, c, phi, Int):
        c, phi, Int = _np(c), _np(phi), _np(Int)
        out = np.zeros(3)
        assert self.l.yho_solve_matrix(_p(c), _p(phi), _p(Int), _p(out)) == 0
        return out

    def cxy_field(self, p, c, phi, solid=None):
        ax, ay = np.zeros(p.nx * p.ny), np.zeros(p.nx * p.ny)
        c, phi = _np(c), _np(phi)
        s = _np(solid, np.uint8) if solid is not None else None
        assert self.l.yho_cxy_field(C.byref(p), _p(ax), _p(ay), _p(c), _p(phi), _p(s)) == 0
        return ax, ay

    def advect_bfecc(self, p, u, v, adv_x, adv_y, solid=None):
        u, v, ax, ay = map(_np, (u, v, adv_x, adv_y))
        uo, vo = np.empty_like(u), np.empty_like(v)
        s = _np(solid, np.uint8) if solid is not None else None
        rc = self.l.yho_advect_bfecc(C.byref(p), _p(u), _p(v), _p(uo), _p(vo), _p(ax), _p(ay), _p(s))
        assert rc == 0, rc
        return uo, vo

    def sapd_sequence(self, p, u_seq, count0=0, stimArea=None, stimulate=False):
        """Run sAPD over consecutive frames (uold=u_seq[k], unew=u_seq[k+1], count=count0+k)."""
        n = p.nx * p.ny
        st = {k: np.zeros(n) for k in ("APD1", "APD2", "sAPD", "dAPD", "back", "front")}
        first = np.zeros(n, dtype=np.uint8)
        sa = _np(stimArea, np.uint8) if stimArea is not None else None
        for k in range(len(u_seq) - 1):
            a, b = _np(u_seq[k]), _np(u_seq[k + 1])
            rc = self.l.yho_sapd(C.byref(p), count0 + k, _p(a), _p(b), _p(st["APD1"]), _p(st["APD2"]),
                                 _p(st["sAPD"]), _p(st["dAPD"]), _p(st["back"]), _p(st["front"]),
                                 _p(first), _p(sa), int(stimulate))
            assert rc == 0, rc
        st["first"] = first
        return st


    def contour(self, p, field1, field2, mode, t=0.0, stimArea=None, thresh=CONTOUR_THRESH, plot=False,
                capacity=None):
        """countour_wrapper (spaceAPD.cu:256-276): points in canonical order (+ raster)."""
        f2 = _np(field2)
        f1 = _np(field1) if field1 is not None else None
        sa = _np(stimArea, np.uint8) if stimArea is not None else None
        cap = capacity if capacity is not None else 2 * f2.size
        vec = np.zeros(cap, dtype=CONTOUR_DTYPE)
        n = C.c_int(0)
        pl = np.full(f2.size, 7, dtype=np.uint8) if plot else None
        rc = self.l.yho_contour(C.byref(p), _p(f1), _p(f2), _p(pl), _p(sa), C.byref(n), _p(vec), cap,
                                _d(t), mode, _d(thresh[0]), _d(thresh[1]), _d(thresh[2]))
        assert rc == 0, rc
        out = vec[: min(n.value, cap)].copy()
        return (out, n.value, pl) if plot else (out, n.value)

    def rgba(self, p, field, cmap, vmin, vmax, lines=None):
        f = _np(field)
        cm = np.ascontiguousarray(cmap, dtype=np.uint32)
        ln = _np(lines, np.uint8) if lines is not None else None
        out = np.zeros(f.size, dtype=np.uint32)
        rc = self.l.yho_rgba(C.byref(p), _p(f), _p(out), _p(cm), len(cm), _d(vmin), _d(vmax), _p(ln))
        assert rc == 0, rc
        return out


class Reference:
    """The reference's own CUDA kernels (oracle/_ref), driven through ref_harness.cu."""

    def __init__(self, nofma=False):
        path = REF_NOFMA_SO if nofma else REF_SO
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.l = C.CDLL(path)
        self.l.yref_rd_run.restype = C.c_float
        self.l.yref_sr_run.restype = C.c_float
        self.l.yref_apd_run.restype = C.c_float
        self.p = None

    def init(self, p):
        assert self.l.yref_init(C.byref(p)) == 0
        self.p = p

    def rd_run(self, u, v, nsteps, solid=None, stim_mouse=False, point=None, mode=0, velTan=False,
               copy_back=True):
        u, v = _np(u).copy(), _np(v).copy()
        p = self.p
        px, py = point if point is not None else (p.nx // 2, p.ny // 2)
        s = _np(solid, np.uint8) if solid is not None else np.ones(p.nx * p.ny, dtype=np.uint8)
        vtu = np.zeros_like(u) if velTan else None
        vtv = np.zeros_like(u) if velTan else None
        ms = self.l.yref_rd_run(_p(u), _p(v), _p(vtu), _p(vtv), _p(s), nsteps, int(stim_mouse), px, py,
                                mode, int(copy_back))
        assert ms >= 0, "reference harness failed"
        return (u, v, vtu, vtv, ms) if velTan else (u, v, ms)

    def sr_run(self, u, v, nsteps):
        """display() symmetry-reduction loop with the reference's own wrappers; -> u, v, c_phi, ms"""
        u, v = _np(u).copy(), _np(v).copy()
        rec = np.zeros((nsteps, 6))
        ms = self.l.yref_sr_run(_p(u), _p(v), nsteps, _p(rec), 1)
        assert ms >= 0
        return u, v, rec, ms

    def apd_run(self, u, v, nsteps, period_it, duration_it, stimArea, mode=0):
        u, v = _np(u).copy(), _np(v).copy()
        sa = _np(stimArea, np.uint8)
        apd = np.zeros(2 * u.size)
        ms = self.l.yref_apd_run(_p(u), _p(v), nsteps, period_it, duration_it, _p(sa), _p(apd), mode)
        assert ms >= 0
        return u, v, apd[:u.size].copy(), apd[u.size:].copy(), ms

    def tip(self, u_present, u_past, t=0.0, algorithm=1, capacity=65536):
        a, b = _np(u_present), _np(u_past)
        vec = np.zeros(capacity, dtype=TIP_DTYPE)
        n = self.l.yref_tip(_p(a), _p(b), _d(t), algorithm, _p(vec), capacity, None)
        assert n >= 0
        return vec[: min(n, capacity)].copy()

    def contour(self, field1, field2, mode, t=0.0, stimArea=None):
        """countour_wrapper with the reference's default thresholds; list in atomicAdd order."""
        f2 = _np(field2)
        f1 = _np(field1) if field1 is not None else None
        sa = _np(stimArea, np.uint8) if stimArea is not None else np.ones(f2.size, dtype=np.uint8)
        cap = 2 * f2.size
        vec = np.zeros(cap, dtype=CONTOUR_DTYPE)
        pl = np.zeros(f2.size, dtype=np.uint8)
        n = self.l.yref_contour(_p(f1), _p(f2), _p(sa), C.c_float(t), mode, _p(vec), cap, _p(pl))
        assert n >= 0
        return vec[:n].copy(), pl

    def slice_trapz(self, u, v, adv_x, adv_y, vtu, vtv, tipx, tipy, count, want_slices=False):
        u, v, ax, ay, vtu, vtv = map(_np, (u, v, adv_x, adv_y, vtu, vtv))
        out = np.zeros(12)
        sl = np.zeros(12 * u.size) if want_slices else None
        rc = self.l.yref_slice_trapz(_p(u), _p(v), _p(ax), _p(ay), _p(vtu), _p(vtv), C.c_float(tipx),
                                     C.c_float(tipy), count, _p(out), _p(sl))
        assert rc == 0
        return (out, sl.reshape(12, -1)) if want_slices else out

    def solve_matrix(self, c, phi, Int):
        c, phi, Int = _np(c), _np(phi), _np(Int).copy()
        out = np.zeros(3)
        assert self.l.yref_solve_matrix(_p(c), _p(phi), _p(Int), _p(out)) == 0
        return out

    def cxy(self, c, phi, solid=None):
        p = self.p
        ax, ay = np.zeros(p.nx * p.ny), np.zeros(p.nx * p.ny)
        c, phi = _np(c), _np(phi)
        s = _np(solid, np.uint8) if solid is not None else np.ones(p.nx * p.ny, dtype=np.uint8)
        assert self.l.yref_cxy(_p(c), _p(phi), _p(s), _p(ax), _p(ay)) == 0
        return ax, ay

    def bfecc(self, u, v, adv_x, adv_y, solid=None, repeats=1):
        p = self.p
        u, v, ax, ay = map(_np, (u, v, adv_x, adv_y))
        uo, vo = np.empty_like(u), np.empty_like(v)
        s = _np(solid, np.uint8) if solid is not None else np.ones(p.nx * p.ny, dtype=np.uint8)
        assert self.l.yref_bfecc(_p(u), _p(v), _p(ax), _p(ay), _p(s), _p(uo), _p(vo), repeats) == 0
        return uo, vo

    def sapd_sequence(self, u_seq, count0=0, stimArea=None, stimulate=False):
        p = self.p
        n = p.nx * p.ny
        seq = _np(np.stack([np.asarray(a).reshape(-1) for a in u_seq]))
        out6 = np.zeros(6 * n)
        first = np.zeros(n, dtype=np.uint8)
        sa = _np(stimArea, np.uint8) if stimArea is not None else np.ones(n, dtype=np.uint8)
        rc = self.l.yref_sapd(_p(seq), len(u_seq), count0, _p(sa), int(stimulate), _p(out6), _p(first))
        assert rc == 0
        names = ("APD1", "APD2", "sAPD", "dAPD", "back", "front")
        st = {k: out6[i * n:(i + 1) * n].copy() for i, k in enumerate(names)}
        st["first"] = first
        return st


_oracle = None


def load():
    global _oracle
    if _oracle is None:
        _oracle = Oracle()
    return _oracle


def have_reference():
    return os.path.exists(REF_SO) and os.path.exists(REF_NOFMA_SO)
