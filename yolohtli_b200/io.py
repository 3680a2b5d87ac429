"""Host-side mirror of include/yolohtli_io.h: the files the reference reads and writes
(saveFiles.cu, printFunctions.cu) plus the lossless snapshot.  Every
function forwards to the C entry point of libyolohtli_b200.so of the same name; arrays are
NumPy (host) arrays.  No GPU is needed for any of these."""
import ctypes as C

import numpy as np

from ._lib import YhRunParams, lib
from .host import CONTOUR_DTYPE, TIP_DTYPE, check

RunParams = YhRunParams


def _b(path):
    return None if path is None else str(path).encode()


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def run_params_default(nx=512, ny=512):
    """parameterSetup (saveFiles.cu:105-231)."""
    rp = YhRunParams()
    check(lib().yh_io_run_params_default(C.byref(rp), nx, ny))
    return rp


def params_write_csv(path, rp):
    """printParameters (printFunctions.cu:266-403)."""
    check(lib().yh_io_params_write_csv(_b(path), C.byref(rp)))


def params_read_csv(path):
    """loadParamValues (saveFiles.cu:540-711)."""
    rp = YhRunParams()
    check(lib().yh_io_params_read_csv(_b(path), C.byref(rp)))
    return rp


def state_write_text(path, u, v):
    """print2D2column (printFunctions.cu:58-79)."""
    u, v = _f64(u), _f64(v)
    ny, nx = u.shape
    check(lib().yh_io_state_write_text(_b(path), _p(u), _p(v), nx, ny))


def state_read_text(path, nx, ny):
    """loadData (saveFiles.cu:508-538)."""
    u, v = np.empty((ny, nx)), np.empty((ny, nx))
    check(lib().yh_io_state_read_text(_b(path), _p(u), _p(v), nx, ny))
    return u, v


def state_write_window(path, u, v, tipx, tipy, offx, offy):
    """print2DSubWindow (printFunctions.cu:81-104); returns the number of cells written."""
    u, v = _f64(u), _f64(v)
    ny, nx = u.shape
    n = C.c_longlong(0)
    check(lib().yh_io_state_write_window(_b(path), _p(u), _p(v), nx, ny, float(tipx), float(tipy),
                                         int(offx), int(offy), C.byref(n)))
    return n.value


def snapshot_write(path, u, v, count=0, physical_time=0.0):
    u, v = _f64(u), _f64(v)
    if u.ndim == 2:
        u, v = u[None], v[None]
    ns, ny, nx = u.shape
    check(lib().yh_io_snapshot_write(_b(path), _p(u), _p(v), nx, ny, ns, int(count), float(physical_time)))


def snapshot_read(path):
    """Returns (u, v, count, physical_time); u, v shaped (n_sims, ny, nx)."""
    nx, ny, ns, cnt, t = C.c_int(), C.c_int(), C.c_int(), C.c_longlong(), C.c_double()
    check(lib().yh_io_snapshot_info(_b(path), C.byref(nx), C.byref(ny), C.byref(ns), C.byref(cnt), C.byref(t)))
    shape = (ns.value, ny.value, nx.value)
    u, v = np.empty(shape), np.empty(shape)
    check(lib().yh_io_snapshot_read(_b(path), _p(u), _p(v), u.size))
    return u, v, cnt.value, t.value


def mask_read(path, nx, ny):
    """holes<N>.dat / cBoundary<N>.dat: value > 0.5 => tissue (main.cu:676-680)."""
    m = np.empty((ny, nx), dtype=np.uint8)
    check(lib().yh_io_mask_read(_b(path), _p(m), m.size))
    return m


def mask_write(path, mask):
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    check(lib().yh_io_mask_write(_b(path), _p(m), m.size))


def domain_objects(rp):
    """domainObjects (main.cu:686-848): returns (intglArea, stimArea, stimulus), shaped (ny, nx)."""
    ny, nx = rp.k.ny, rp.k.nx
    ia, sa = np.zeros((ny, nx), dtype=np.uint8), np.zeros((ny, nx), dtype=np.uint8)
    st = np.zeros((ny, nx))
    check(lib().yh_io_domain_objects(C.byref(rp), _p(ia), _p(sa), _p(st)))
    return ia, sa, st


def tips_append(path_points, path_counts, tips, first=False):
    """printTip (printFunctions.cu:149-197)."""
    t = np.ascontiguousarray(tips, dtype=TIP_DTYPE)
    check(lib().yh_io_tips_append(_b(path_points), _b(path_counts), _p(t), len(t), int(first)))


def contour_append(path_points, path_counts, pts, first=False):
    """printContour (printFunctions.cu:199-247)."""
    t = np.ascontiguousarray(pts, dtype=CONTOUR_DTYPE)
    check(lib().yh_io_contour_append(_b(path_points), _b(path_counts), _p(t), len(t), int(first)))


def sym_write(path, c_phi):
    """printSym (printFunctions.cu:249-264); c_phi: (nsteps, 6) = c then phi."""
    a = _f64(c_phi).reshape(-1, 6)
    check(lib().yh_io_sym_write(_b(path), _p(a), len(a)))


def series_write(path, e0, e1, dt, it_per_frame):
    """printVoltageInTime (printFunctions.cu:106-126)."""
    e0, e1 = _f64(e0), _f64(e1)
    check(lib().yh_io_series_write(_b(path), _p(e0), _p(e1), len(e0), float(dt), int(it_per_frame)))


def contour_length_write(path, lengths, dt, it_per_frame):
    """printContourLengthInTime (printFunctions.cu:128-147)."""
    a = _f64(lengths)
    check(lib().yh_io_contour_length_write(_b(path), _p(a), len(a), float(dt), int(it_per_frame)))


def reconstruct_tip(tip_x, tip_y, c_phi, dx, dy):
    """Original-frame tip path from the symmetry-reduced one (DATA/processSymmetry.m:68-89)."""
    x = np.ascontiguousarray(tip_x, dtype=np.float32)
    y = np.ascontiguousarray(tip_y, dtype=np.float32)
    a = _f64(c_phi).reshape(-1, 6)
    X, Y = np.empty(len(x)), np.empty(len(x))
    check(lib().yh_io_reconstruct_tip(_p(x), _p(y), _p(a), len(x), float(dx), float(dy), _p(X), _p(Y)))
    return X, Y


def cmap_read(path=None, capacity=4096):
    """loadcmap (main.cu:1434-1470); path=None: built-in ramp of `capacity` colours."""
    cm = np.zeros(capacity, dtype=np.uint32)
    n = C.c_int(0)
    check(lib().yh_io_cmap_read(_b(path), _p(cm), capacity, C.byref(n)))
    return cm[: n.value].copy()


def frame_write_ppm(path, rgba):
    a = np.ascontiguousarray(rgba, dtype=np.uint32)
    ny, nx = a.shape
    check(lib().yh_io_frame_write_ppm(_b(path), _p(a), nx, ny))
