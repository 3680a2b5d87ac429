"""Host-side mirror of the reference's launch interface, over the C ABI.

Function names follow the reference's wrappers (hostPrototypes.h:22-54): rd_step,
tip_track, slice_fields, trapz, cxy_field, advect_bfecc, solve_matrix, sapd, probe.  Arrays
are torch CUDA tensors (float64 fields, uint8 masks); torch is only the owner of device
memory and streams here -- every computation is a kernel of libyolohtli_b200.so.
"""
import ctypes as C

import numpy as np

from ._lib import YhParams, YhTip, YolohtliError, lib

Params = YhParams
TIP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("vx", "<f4"), ("vy", "<f4"), ("t", "<f4")])
TIPVECSIZE = 500000


def check(rc):
    if rc != 0:
        msg = lib().yh_last_error()
        raise YolohtliError(f"libyolohtli_b200 status {rc}: {msg.decode() if msg else ''}")


def default_params(nx=512, ny=512, reduce_sym=False, scale_L=False, **over):
    """parameterSetup() defaults (saveFiles.cu:105-231) + main.cu:148-158, then overrides."""
    p = YhParams()
    check(lib().yh_params_default(C.byref(p), nx, ny, int(reduce_sym), int(scale_L)))
    for k, v in over.items():
        if not hasattr(p, k):
            raise AttributeError(f"yh_params has no field {k}")
        setattr(p, k, v)
    return p


def _ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f64(t, n=None):
    import torch
    assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous(), "need contiguous CUDA float64"
    if n is not None:
        assert t.numel() >= n
    return t


def rd_step(p, u_in, v_in, u_out, v_out, velTan=None, solid=None, stim_mouse=False, point=None,
            rows=None):
    """reactionDiffusion_wrapper (reactionDiffusion.cu:568-577)."""
    n = p.nx * p.ny
    _f64(u_in, n), _f64(v_in, n), _f64(u_out, n), _f64(v_out, n)
    px, py = point if point is not None else (p.nx // 2, p.ny_global // 2)
    r0, r1 = rows if rows is not None else (0, p.ny)
    vtu, vtv = velTan if velTan is not None else (None, None)
    check(lib().yh_rd_step(C.byref(p), _ptr(u_in), _ptr(v_in), _ptr(u_out), _ptr(v_out), _ptr(vtu),
                           _ptr(vtv), _ptr(solid), int(stim_mouse), px, py, r0, r1, _stream()))


RD_INPUT_CANONICAL = 1
RD_SOLID_IS_PATTERNS = 2


def rd_mask_patterns(p, solid, patterns):
    """Per-cell mask patterns of the temporally blocked masked Euler kernel (computed once for a
    fixed mask; pass as `solid` to rd_advance with flags |= RD_SOLID_IS_PATTERNS)."""
    check(lib().yh_rd_mask_patterns(C.byref(p), _ptr(solid), _ptr(patterns), _stream()))
    return patterns


def rd_advance(p, nsteps, uA, vA, uB, vB, tb_steps=0, solid=None, stim_mouse=False, point=None,
               rows=None, flags=0):
    """nsteps x {reactionDiffusion_wrapper; swapSoA} (main.cu:879-882).  Returns (u, v) tensors
    holding the result (either the A or the B pair)."""
    px, py = point if point is not None else (p.nx // 2, p.ny_global // 2)
    r0, r1 = rows if rows is not None else (0, p.ny)
    inB = C.c_int(0)
    check(lib().yh_rd_advance(C.byref(p), nsteps, tb_steps, flags, _ptr(uA), _ptr(vA), _ptr(uB), _ptr(vB),
                              _ptr(solid), int(stim_mouse), px, py, r0, r1, C.byref(inB), _stream()))
    return (uB, vB) if inB.value else (uA, vA)


def tip_track(p, u_past, u_present, tip_count, tip_vector, tip_plot=None, t=0.0, algorithm=None,
              capacity=TIPVECSIZE):
    """tip_wrapper (tipTracker.cu:569-611); tip_vector: CUDA uint8 tensor of capacity*20 bytes."""
    alg = algorithm if algorithm is not None else p.tipAlgorithm
    check(lib().yh_tip_track(C.byref(p), _ptr(u_past), _ptr(u_present), _ptr(tip_plot),
                             _ptr(tip_count), _ptr(tip_vector), capacity, float(t), alg, _stream()))


def tip_track_rows(p, u_past, u_present, tip_count, tip_vector, rows, t=0.0, algorithm=None,
                   capacity=TIPVECSIZE):
    """Row-slab form of tip_track: cells of local rows [rows[0], rows[1]), global coordinates."""
    alg = algorithm if algorithm is not None else p.tipAlgorithm
    check(lib().yh_tip_track_rows(C.byref(p), _ptr(u_past), _ptr(u_present), None, _ptr(tip_count),
                                  _ptr(tip_vector), capacity, float(t), alg, rows[0], rows[1], _stream()))


def tips_to_numpy(tip_count, tip_vector):
    n = int(tip_count.item())
    raw = tip_vector[: n * 20].cpu().numpy().tobytes()
    return np.frombuffer(raw, dtype=TIP_DTYPE).copy()


def _parr(ts):
    arr = (C.c_void_p * 6)()
    for i, t in enumerate(ts):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


def slice_fields(p, u, v, slice6, slice06, adv_x, adv_y, reduce_sym=True, reduce_sym_start=True,
                 scheme=2, tip_count=None, tip_vector=None, count=0):
    """slice_wrapper (symmetryReduction.cu:312-322); slice6/slice06: lists ux,uy,ut,vx,vy,vt."""
    check(lib().yh_slice(C.byref(p), _ptr(u), _ptr(v), _parr(slice6), _parr(slice06),
                         int(reduce_sym), int(reduce_sym_start), _ptr(adv_x), _ptr(adv_y), scheme,
                         _ptr(tip_count), _ptr(tip_vector), count, _stream()))


def trapz(p, slice6, slice06, velTan_u, velTan_v, tip_count=None, tip_vector=None, count=0):
    """trapz_wrapper (integralTrapz.cu:85-184) -> numpy[12] (host)."""
    out = (C.c_double * 12)()
    check(lib().yh_trapz(C.byref(p), _parr(slice6), _parr(slice06), _ptr(velTan_u), _ptr(velTan_v),
                         out, _ptr(tip_count), _ptr(tip_vector), count, _stream()))
    return np.array(out[:], dtype=np.float64)


def sr_integrals(p, u, v, velTan_u, velTan_v, adv_x, adv_y, tip_count=None, tip_vector=None, count=0):
    out = (C.c_double * 12)()
    check(lib().yh_sr_integrals(C.byref(p), _ptr(u), _ptr(v), _ptr(velTan_u), _ptr(velTan_v),
                                _ptr(adv_x), _ptr(adv_y), out, _ptr(tip_count), _ptr(tip_vector),
                                count, _stream()))
    return np.array(out[:], dtype=np.float64)


def sr_disc_slots(p):
    return lib().yh_sr_disc_slots(C.byref(p))


def sr_integral_rows(p, u, v, velTan_u, velTan_v, adv_x, adv_y, centre, rows, rows_d):
    """Row-slab form of sr_integrals, part 1: row sums of the owned disc rows -> rows_d (device)."""
    assert rows_d.numel() >= 12 * sr_disc_slots(p)
    check(lib().yh_sr_integral_rows(C.byref(p), _ptr(u), _ptr(v), _ptr(velTan_u), _ptr(velTan_v),
                                    _ptr(adv_x), _ptr(adv_y), float(centre[0]), float(centre[1]),
                                    rows[0], rows[1], _ptr(_f64(rows_d)), _stream()))


def sr_integrals_close(p, rows_d):
    """Part 2: the 12 integrals (host) from row sums assembled over all slabs."""
    out = (C.c_double * 12)()
    check(lib().yh_sr_integrals_close(C.byref(p), _ptr(_f64(rows_d)), out, _stream()))
    return np.array(out[:], dtype=np.float64)


def _d3(a):
    return (C.c_double * 3)(*[float(x) for x in a])


def solve_matrix(c, phi, Int):
    """solve_matrix (symmetryReduction.cu:329-420), host."""
    out = (C.c_double * 3)()
    check(lib().yh_solve_matrix(_d3(c), _d3(phi), (C.c_double * 12)(*[float(x) for x in Int]), out))
    return np.array(out[:])


def cxy_field(p, adv_x, adv_y, c, phi, solid=None):
    check(lib().yh_cxy_field(C.byref(p), _ptr(adv_x), _ptr(adv_y), _d3(c), _d3(phi), _ptr(solid),
                             _stream()))


def advect_bfecc(p, u_in, v_in, u_out, v_out, adv_x, adv_y, solid=None):
    """advFDBFECC_wrapper (advFDBFECC.cu:355-362)."""
    check(lib().yh_advect_bfecc(C.byref(p), _ptr(u_in), _ptr(v_in), _ptr(u_out), _ptr(v_out),
                                _ptr(adv_x), _ptr(adv_y), _ptr(solid), _stream()))


def advect_bfecc_cphi(p, u_in, v_in, u_out, v_out, c, phi, adv_x=None, adv_y=None, solid=None):
    check(lib().yh_advect_bfecc_cphi(C.byref(p), _ptr(u_in), _ptr(v_in), _ptr(u_out), _ptr(v_out),
                                     _d3(c), _d3(phi), _ptr(adv_x), _ptr(adv_y), _ptr(solid),
                                     _stream()))


def advect_bfecc_cphi_rows(p, u_in, v_in, u_out, v_out, c, phi, rows, adv_x=None, adv_y=None, solid=None):
    check(lib().yh_advect_bfecc_cphi_rows(C.byref(p), _ptr(u_in), _ptr(v_in), _ptr(u_out), _ptr(v_out),
                                          _d3(c), _d3(phi), _ptr(adv_x), _ptr(adv_y), _ptr(solid),
                                          rows[0], rows[1], _stream()))


def sapd(p, count, uold, unew, APD1, APD2, sAPD, dAPD, back, front, first, stimArea=None,
         stimulate=False):
    """sAPD_wrapper (spaceAPD.cu:376-384)."""
    check(lib().yh_sapd(C.byref(p), count, _ptr(uold), _ptr(unew), _ptr(APD1), _ptr(APD2), _ptr(sAPD),
                        _ptr(dAPD), _ptr(back), _ptr(front), _ptr(first), _ptr(stimArea),
                        int(stimulate), _stream()))


CONTOUR_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("t", "<f4")])
CONTOUR_THRESH = (0.8, 0.85, 0.7)   # contourThresh1..3, saveFiles.cu:215-217


def contour(p, field1, field2, contour_count, contour_vector, mode, contour_plot=None, stimArea=None,
            t=0.0, thresh=CONTOUR_THRESH, capacity=None):
    """countour_wrapper (spaceAPD.cu:256-276).  contour_vector: CUDA uint8 tensor of capacity*12
    bytes; contour_count: CUDA int32 tensor of one element."""
    cap = capacity if capacity is not None else contour_vector.numel() // 12
    check(lib().yh_contour(C.byref(p), _ptr(field1), _ptr(field2), _ptr(contour_plot), _ptr(stimArea),
                           _ptr(contour_count), _ptr(contour_vector), cap, float(t), mode,
                           float(thresh[0]), float(thresh[1]), float(thresh[2]), _stream()))


def contour_to_numpy(contour_count, contour_vector):
    n = int(contour_count.item())
    m = min(n, contour_vector.numel() // 12)
    raw = contour_vector[: m * 12].cpu().numpy().tobytes()
    return np.frombuffer(raw, dtype=CONTOUR_DTYPE).copy(), n


def rgba(p, field, plot_rgba, cmap_rgba, min_var=-0.1, max_var=1.1, lines=None):
    """get_rgba_wrapper (main.cu:1633-1641); plot_rgba / cmap_rgba: CUDA int32 tensors holding the
    packed 0xAABBGGRR words."""
    check(lib().yh_rgba(C.byref(p), _ptr(field), _ptr(plot_rgba), _ptr(cmap_rgba), cmap_rgba.numel(),
                        float(min_var), float(max_var), _ptr(lines), _stream()))


def probe(p, u, v, pt_d, x, y):
    """singleCell_wrapper (singleCell.cu:22-30) without the blocking copy."""
    check(lib().yh_probe(C.byref(p), _ptr(u), _ptr(v), _ptr(pt_d), x, y, None, _stream()))


class Sim:
    """Headless driver (yh_sim_*): the display() loop of main.cu:862-1043 without GL."""

    def __init__(self, p, n_sims=1, device=0):
        self.p = p
        self.n_sims = n_sims
        self._h = C.c_void_p()
        check(lib().yh_sim_create(C.byref(self._h), C.byref(p), n_sims, device))
        self.shape = (n_sims, p.ny, p.nx)

    def close(self):
        if self._h:
            lib().yh_sim_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def cross_field_ic(self):
        check(lib().yh_sim_cross_field_ic(self._h))

    def set_state(self, u, v):
        u = np.ascontiguousarray(u, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert u.size == v.size == int(np.prod(self.shape))
        check(lib().yh_sim_set_state(self._h, u.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p)))

    def get_state(self, out=None):
        """State of every sheet; out = (u, v) host arrays to fill (e.g. pinned), else fresh ones."""
        u, v = out if out is not None else (np.empty(self.shape, dtype=np.float64), np.empty(self.shape, dtype=np.float64))
        assert u.size == v.size == int(np.prod(self.shape)) and u.dtype == v.dtype == np.float64
        check(lib().yh_sim_get_state(self._h, u.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p)))
        return u, v

    def set_solid(self, solid):
        s = np.ascontiguousarray(solid, dtype=np.uint8)
        assert s.size == self.p.nx * self.p.ny
        check(lib().yh_sim_set_solid(self._h, s.ctypes.data_as(C.c_void_p)))

    def set_point(self, x, y):
        check(lib().yh_sim_set_point(self._h, x, y))

    def set_pacing(self, period_it, duration_it):
        if period_it is None:
            check(lib().yh_sim_set_pacing(self._h, None, 0))
            return
        a = np.ascontiguousarray(period_it, dtype=np.int32)
        assert a.size == self.n_sims
        check(lib().yh_sim_set_pacing(self._h, a.ctypes.data_as(C.c_void_p), int(duration_it)))

    def run(self, nsteps, tb_steps=0, trace=False):
        tr = None
        ptr = None
        if trace:
            tr = np.empty((nsteps, self.n_sims, 2), dtype=np.float64)
            ptr = tr.ctypes.data_as(C.c_void_p)
        check(lib().yh_sim_run(self._h, nsteps, tb_steps, ptr))
        return tr

    def run_host(self, u_in, v_in, u_out, v_out, nsteps, tb_steps=0):
        """H2D + nsteps + D2H in one call; arguments are host pointers (ints) or numpy arrays."""
        def hp(a):
            return C.c_void_p(a) if isinstance(a, int) else a.ctypes.data_as(C.c_void_p)
        check(lib().yh_sim_run_host(self._h, hp(u_in), hp(v_in), hp(u_out), hp(v_out), nsteps, tb_steps))

    def run_sr(self, nsteps, record=True):
        """Symmetry-reduction steps (main.cu:894-954); returns [nsteps, 6] = (c, phi) per step."""
        out = np.zeros((nsteps, 6), dtype=np.float64) if record else None
        check(lib().yh_sim_run_sr(self._h, nsteps, out.ctypes.data_as(C.c_void_p) if record else None))
        return out

    def run_sr_device(self, nsteps, record=True):
        """Same steps with the drift solve resident on the device (no host round trip per step)."""
        out = np.zeros((nsteps, 6), dtype=np.float64) if record else None
        check(lib().yh_sim_run_sr_device(self._h, nsteps, out.ctypes.data_as(C.c_void_p) if record else None))
        return out

    def run_apd(self, nsteps, stim_area=None):
        """contourMode == 1 loop: RD + sAPD every step (main.cu:1035) for every sheet."""
        ptr = None
        if stim_area is not None:
            sa = np.ascontiguousarray(stim_area, dtype=np.uint8)
            assert sa.size == self.p.nx * self.p.ny
            ptr = sa.ctypes.data_as(C.c_void_p)
        check(lib().yh_sim_run_apd(self._h, nsteps, ptr))

    def get_apd(self, out=None):
        a, b = out if out is not None else (np.empty(self.shape, dtype=np.float64), np.empty(self.shape, dtype=np.float64))
        assert a.size == b.size == int(np.prod(self.shape)) and a.dtype == b.dtype == np.float64
        check(lib().yh_sim_get_apd(self._h, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)))
        return a, b

    def sr_state(self):
        c, phi = (C.c_double * 3)(), (C.c_double * 3)()
        check(lib().yh_sim_sr_state(self._h, c, phi, 0))
        return np.array(c[:]), np.array(phi[:])

    def tips(self, capacity=4096):
        buf = np.zeros(capacity, dtype=TIP_DTYPE)
        n = C.c_int(0)
        check(lib().yh_sim_tips(self._h, buf.ctypes.data_as(C.c_void_p), capacity, C.byref(n)))
        return buf[: min(n.value, capacity)].copy()

    @property
    def count(self):
        return lib().yh_sim_count(self._h)
