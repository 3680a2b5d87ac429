"""ctypes binding of libyolohtli_b200.so (include/yolohtli_abi.h).  Fails loudly when the
CUDA library has not been built -- there is no Python/NumPy fallback."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("YH_LIB_PATH") or os.path.join(HERE, "lib", "libyolohtli_b200.so")   # override: experiments only
SHIM_PATH = os.path.join(HERE, "lib", "libyolohtli_shim.so")


class YolohtliError(RuntimeError):
    pass


class YhParams(C.Structure):
    """struct yh_params of include/yolohtli_abi.h (same order, same types)."""
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("ny_global", C.c_int32), ("jg0", C.c_int32),
        ("solidSwitch", C.c_int32), ("neumannBC", C.c_int32), ("gateDiff", C.c_int32),
        ("anisotropy", C.c_int32), ("lap4", C.c_int32), ("timeIntOrder", C.c_int32),
        ("tipGrad", C.c_int32), ("tipAlgorithm", C.c_int32),
        ("tipOffsetX", C.c_int32), ("tipOffsetY", C.c_int32),
        ("tipx0", C.c_float), ("tipy0", C.c_float),
        ("dt", C.c_double), ("hx", C.c_double), ("hy", C.c_double), ("Lx", C.c_double),
        ("Ly", C.c_double),
        ("rx", C.c_double), ("ry", C.c_double), ("rxy", C.c_double), ("rbx", C.c_double),
        ("rby", C.c_double), ("rscale", C.c_double),
        ("qx4", C.c_double), ("qy4", C.c_double), ("fx4", C.c_double), ("fy4", C.c_double),
        ("invdx", C.c_double), ("invdy", C.c_double),
        ("tc", C.c_double), ("alpha", C.c_double), ("beta", C.c_double), ("gamma", C.c_double),
        ("delta", C.c_double), ("eps", C.c_double), ("mu", C.c_double), ("theta", C.c_double),
        ("boundaryVal", C.c_double), ("Uth", C.c_double),
    ]

    def copy(self):
        q = YhParams()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(YhParams))
        return q


class YhTip(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("vx", C.c_float), ("vy", C.c_float),
                ("t", C.c_float)]


class YhContourPt(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("t", C.c_float)]


class YhRunParams(C.Structure):
    """struct yh_run_params of include/yolohtli_io.h (same order, same types)."""
    _fields_ = [
        ("k", YhParams),
        ("read_path", C.c_char * 200), ("results_path", C.c_char * 200),
        ("saveEveryIt", C.c_int32), ("plotTip", C.c_int32), ("recordTip", C.c_int32),
        ("plotContour", C.c_int32), ("recordContour", C.c_int32), ("stimulate", C.c_int32),
        ("plotTimeSeries", C.c_int32), ("recordTimeSeries", C.c_int32), ("reduceSym", C.c_int32),
        ("contourMode", C.c_int32), ("clock", C.c_int32), ("counterclock", C.c_int32),
        ("diff_par", C.c_double), ("diff_per", C.c_double), ("degrad", C.c_double),
        ("Dxx", C.c_double), ("Dyy", C.c_double), ("Dxy", C.c_double),
        ("physicalTimeLim", C.c_double), ("startRecTime", C.c_double),
        ("eSize", C.c_int32), ("point_x", C.c_int32), ("point_y", C.c_int32),
        ("stimPeriod", C.c_double), ("stimMag", C.c_double), ("stimDuration", C.c_double),
        ("fibThreshold", C.c_double),
        ("fibTerminated", C.c_int32), ("leapShocks", C.c_int32), ("nc", C.c_int32),
        ("stcx", C.c_double), ("stcy", C.c_double), ("rdomStim", C.c_double),
        ("rdomAPD", C.c_double), ("rdomTrapz", C.c_double),
        ("itPerFrame", C.c_int32),
        ("sample", C.c_double),
        ("minVarColor", C.c_double), ("maxVarColor", C.c_double),
        ("wnx", C.c_int32), ("wny", C.c_int32),
        ("uMin", C.c_double), ("uMax", C.c_double), ("vMin", C.c_double), ("vMax", C.c_double),
        ("tipx", C.c_double), ("tipy", C.c_double),
        ("contourThresh1", C.c_double), ("contourThresh2", C.c_double),
        ("contourThresh3", C.c_double),
    ]


_P = C.POINTER(YhParams)
_RP = C.POINTER(YhRunParams)
_s = C.c_char_p
_ll = C.c_longlong
_vp = C.c_void_p
_i = C.c_int
_d = C.c_double

# name -> (restype, argtypes); every symbol include/yolohtli_abi.h declares
SIGNATURES = {
    "yh_abi_version": (_i, []),
    "yh_last_error": (C.c_char_p, []),
    "yh_device_count": (_i, []),
    "yh_release_workspace": (_i, []),
    "yh_set_arithmetic": (_i, [_i]),
    "yh_get_arithmetic": (_i, []),
    "yh_params_default": (_i, [_P, _i, _i, _i, _i]),
    "yh_params_derive": (_i, [_P, _d, _d, _d]),
    "yh_rd_step": (_i, [_P, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "yh_rd_advance": (_i, [_P, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i,
                           C.POINTER(_i), _vp]),
    "yh_rd_mask_patterns": (_i, [_P, _vp, _vp, _vp]),
    "yh_tip_track": (_i, [_P, _vp, _vp, _vp, _vp, _vp, _i, _d, _i, _vp]),
    "yh_slice": (_i, [_P, _vp, _vp, C.POINTER(_vp), C.POINTER(_vp), _i, _i, _vp, _vp, _i, _vp,
                      _vp, _i, _vp]),
    "yh_trapz": (_i, [_P, C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, C.POINTER(_d), _vp, _vp, _i,
                      _vp]),
    "yh_sr_integrals": (_i, [_P, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_d), _vp, _vp, _i, _vp]),
    "yh_solve_matrix": (_i, [C.POINTER(_d), C.POINTER(_d), C.POINTER(_d), C.POINTER(_d)]),
    "yh_cxy_field": (_i, [_P, _vp, _vp, C.POINTER(_d), C.POINTER(_d), _vp, _vp]),
    "yh_advect_bfecc": (_i, [_P, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "yh_advect_bfecc_cphi": (_i, [_P, _vp, _vp, _vp, _vp, C.POINTER(_d), C.POINTER(_d), _vp, _vp,
                                  _vp, _vp]),
    "yh_sapd": (_i, [_P, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "yh_probe": (_i, [_P, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "yh_contour": (_i, [_P, _vp, _vp, _vp, _vp, _vp, _vp, _i, _d, _i, _d, _d, _d, _vp]),
    "yh_rgba": (_i, [_P, _vp, _vp, _vp, _i, _d, _d, _vp, _vp]),
    "yh_sim_create": (_i, [C.POINTER(_vp), _P, _i, _i]),
    "yh_sim_destroy": (_i, [_vp]),
    "yh_sim_set_state": (_i, [_vp, _vp, _vp]),
    "yh_sim_get_state": (_i, [_vp, _vp, _vp]),
    "yh_sim_set_solid": (_i, [_vp, _vp]),
    "yh_sim_cross_field_ic": (_i, [_vp]),
    "yh_sim_set_point": (_i, [_vp, _i, _i]),
    "yh_sim_run": (_i, [_vp, _i, _i, _vp]),
    "yh_sim_set_pacing": (_i, [_vp, _vp, _i]),
    "yh_sim_run_host": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i]),
    "yh_sim_tips": (_i, [_vp, _vp, _i, C.POINTER(_i)]),
    "yh_sim_run_sr": (_i, [_vp, _i, _vp]),
    "yh_sim_run_sr_device": (_i, [_vp, _i, _vp]),
    "yh_sim_run_apd": (_i, [_vp, _i, _vp]),
    "yh_sim_get_apd": (_i, [_vp, _vp, _vp]),
    "yh_sim_sr_state": (_i, [_vp, C.POINTER(_d), C.POINTER(_d), _i]),
    "yh_tip_track_rows": (_i, [_P, _vp, _vp, _vp, _vp, _vp, _i, _d, _i, _i, _i, _vp]),
    "yh_sr_disc_slots": (_i, [_P]),
    "yh_sr_integral_rows": (_i, [_P, _vp, _vp, _vp, _vp, _vp, _vp, C.c_float, C.c_float, _i, _i, _vp, _vp]),
    "yh_sr_integrals_close": (_i, [_P, _vp, C.POINTER(_d), _vp]),
    "yh_advect_bfecc_cphi_rows": (_i, [_P, _vp, _vp, _vp, _vp, C.POINTER(_d), C.POINTER(_d), _vp, _vp,
                                       _vp, _i, _i, _vp]),
    "yh_flag_set": (_i, [_vp, _i, _vp]),
    "yh_flag_wait": (_i, [_vp, _i, _vp, _vp]),
    "yh_memcpy_async": (_i, [_vp, _vp, C.c_size_t, _vp]),
    "yh_enable_peer_access": (_i, [_i]),
    "yh_sim_count": (_i, [_vp]),
    "yh_sim_device_u": (_vp, [_vp]),
    "yh_sim_device_v": (_vp, [_vp]),
}

# include/yolohtli_slab.h -- multi-GPU row-slab driver (csrc/slab.cu)
_ull = C.c_ulonglong
SIGNATURES.update({
    "yh_slab_partition": (_i, [_i, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "yh_slab_create": (_i, [C.POINTER(_vp), _P, _i, _i, _i, _i]),
    "yh_slab_destroy": (_i, [_vp]),
    "yh_slab_layout": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "yh_slab_export": (_i, [_vp, _vp]),
    "yh_slab_connect": (_i, [_vp, _vp, _vp]),
    "yh_slab_connect_local": (_i, [_vp, _vp, _vp]),
    "yh_slab_set_state": (_i, [_vp, _vp, _vp, _i]),
    "yh_slab_get_state": (_i, [_vp, _vp, _vp]),
    "yh_slab_set_solid": (_i, [_vp, _vp]),
    "yh_slab_device_u": (_vp, [_vp]),
    "yh_slab_device_v": (_vp, [_vp]),
    "yh_slab_stream": (_vp, [_vp]),
    "yh_slab_advance": (_i, [_vp, _i, _i]),
    "yh_slab_sync": (_i, [_vp]),
    "yh_slab_checksum": (_i, [_vp, C.POINTER(_ull), C.POINTER(_ull)]),
    "yh_slab_run_host": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i]),
    "yh_slab_group_create": (_i, [C.POINTER(_vp), _P, _i, C.POINTER(_i), _i]),
    "yh_slab_group_destroy": (_i, [_vp]),
    "yh_slab_group_member": (_vp, [_vp, _i]),
    "yh_slab_group_set_state": (_i, [_vp, _vp, _vp]),
    "yh_slab_group_get_state": (_i, [_vp, _vp, _vp]),
    "yh_slab_group_set_solid": (_i, [_vp, _vp]),
    "yh_slab_group_advance": (_i, [_vp, _i, _i]),
    "yh_slab_group_sync": (_i, [_vp]),
    "yh_slab_group_run_host": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i]),
    "yh_rd_tile_march_tiling": (_i, [_i, _i, _i, _i, _vp]),
    "yh_slab_pipeline_levels": (_i, [_vp, _i, _i]),
    "yh_slab_pipeline_plan": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "yh_slab_pipeline_region": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "yh_slab_group_advance_sr": (_i, [_vp, _i, _vp]),
    "yh_slab_group_sr_state": (_i, [_vp, _vp, _vp, _i]),
})
SLAB_HANDLE_BYTES = 256

# include/yolohtli_io.h -- host-side data formats (no GPU needed)
SIGNATURES.update({
    "yh_io_run_params_default": (_i, [_RP, _i, _i]),
    "yh_io_params_write_csv": (_i, [_s, _RP]),
    "yh_io_params_read_csv": (_i, [_s, _RP]),
    "yh_io_state_write_text": (_i, [_s, _vp, _vp, _i, _i]),
    "yh_io_state_read_text": (_i, [_s, _vp, _vp, _i, _i]),
    "yh_io_state_write_window": (_i, [_s, _vp, _vp, _i, _i, _d, _d, _i, _i, C.POINTER(_ll)]),
    "yh_io_snapshot_write": (_i, [_s, _vp, _vp, _i, _i, _i, _ll, _d]),
    "yh_io_snapshot_info": (_i, [_s, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_ll),
                                 C.POINTER(_d)]),
    "yh_io_snapshot_read": (_i, [_s, _vp, _vp, _ll]),
    "yh_io_mask_read": (_i, [_s, _vp, _ll]),
    "yh_io_mask_write": (_i, [_s, _vp, _ll]),
    "yh_io_domain_objects": (_i, [_RP, _vp, _vp, _vp]),
    "yh_io_tips_append": (_i, [_s, _s, _vp, _i, _i]),
    "yh_io_contour_append": (_i, [_s, _s, _vp, _i, _i]),
    "yh_io_sym_write": (_i, [_s, _vp, _i]),
    "yh_io_series_write": (_i, [_s, _vp, _vp, _i, _d, _i]),
    "yh_io_contour_length_write": (_i, [_s, _vp, _i, _d, _i]),
    "yh_io_reconstruct_tip": (_i, [_vp, _vp, _vp, _i, _d, _d, _vp, _vp]),
    "yh_io_cmap_read": (_i, [_s, _vp, _i, C.POINTER(_i)]),
    "yh_io_frame_write_ppm": (_i, [_s, _vp, _i, _i]),
})

_lib = None


def load_library(path=None):
    """dlopen the CUDA library and type every entry point.  Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise YolohtliError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  yolohtli_b200 has no CPU fallback.")
    lib_ = C.CDLL(p, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib_, name)   # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib_
    return lib_


def lib():
    return load_library()
