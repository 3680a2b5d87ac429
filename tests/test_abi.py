"""CPU: the C-ABI library loads, exports every symbol include/yolohtli_abi.h declares, agrees
with the oracle on the parameter block, and refuses to compute without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    syms = set()
    for h in ("yolohtli_abi.h", "yolohtli_slab.h", "yolohtli_io.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        syms |= set(re.findall(r"\b(yh_[a-z0-9_]+)\s*\(", txt))
    return sorted(syms)


def test_library_exports_every_declared_symbol(yh):
    from yolohtli_b200 import _lib
    l = yh.lib()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(l, s), f"{s} declared in include/*.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert l.yh_abi_version() == 1


def test_shim_exports_reference_signatures():
    from yolohtli_b200 import _lib
    import subprocess
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.SHIM_PATH]).decode()
    # mangled names of hostPrototypes.h:22-57 (plain C++ linkage, structs by value)
    for frag in ("reactionDiffusion_wrapper", "tip_wrapper", "slice_wrapper", "Cxy_field_wrapper",
                 "advFDBFECC_wrapper", "solve_matrix", "trapz_wrapper", "singleCell_wrapper",
                 "sAPD_wrapper", "countour_wrapper", "get_rgba_wrapper", "swapSoA", "yh_shim_configure",
                 "yh_shim_set_contour_thresholds", "yh_shim_set_color_range"):
        assert frag in out, frag
    assert "_Z25reactionDiffusion_wrapperm4dim3S_8stateVarS0_S0_S0_bPbbPdb4int2" in out


def test_params_default_matches_oracle(yh, oracle):
    for nx, ny, rs, sc in [(512, 512, 0, 0), (512, 512, 1, 0), (1024, 1024, 0, 1), (500, 300, 0, 1)]:
        a = yh.default_params(nx, ny, bool(rs), bool(sc))
        b = oracle.params_default(nx, ny, bool(rs), bool(sc))
        assert bytes(a) == bytes(b)
    p = yh.default_params()
    # saveFiles.cu:124-170 / SURVEY appendix A
    assert (p.lap4, p.timeIntOrder, p.gateDiff, p.neumannBC, p.solidSwitch) == (4, 4, 1, 1, 0)
    assert abs(p.hx - 12.0 / 511.0) < 1e-15 and abs(p.rx - 0.02 * 0.001 / p.hx ** 2) < 1e-15
    assert abs(p.invdx - 0.5 / p.hx) < 1e-12 and p.Uth == 0.7 and p.tipOffsetX == 160


def test_solve_matrix_host(yh, oracle):
    rng = np.random.default_rng(3)
    for _ in range(20):
        Int = rng.normal(size=12)
        phi = rng.normal(size=3)
        got = yh.host.solve_matrix([0, 0, 0], phi, Int)
        want = oracle.solve_matrix([0, 0, 0], phi, Int)
        assert np.array_equal(got, want)
        # it solves A c = d with the first two columns rotated by phi.t (symmetryReduction.cu:390-392)
        cs, sn = np.cos(phi[2]), np.sin(phi[2])
        M = Int[:9].reshape(3, 3)
        A = np.stack([M[:, 0] * cs + M[:, 1] * sn, M[:, 1] * cs - M[:, 0] * sn, M[:, 2]], axis=1)
        assert np.allclose(A @ got, Int[9:], rtol=1e-8, atol=1e-8)


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback(yh):
    l = yh.lib()
    assert l.yh_device_count() == 0
    p = yh.default_params(64, 64)
    buf = np.zeros(64 * 64)
    ptr = buf.ctypes.data_as(C.c_void_p)
    rc = l.yh_rd_step(C.byref(p), ptr, ptr, ptr, ptr, None, None, None, 0, 0, 0, 0, 64, None)
    assert rc == -5   # YH_ERR_NO_DEVICE
    assert b"no CPU fallback" in l.yh_last_error()
    h = C.c_void_p()
    assert l.yh_sim_create(C.byref(h), C.byref(p), 1, 0) == -5


def test_missing_library_fails_loudly(yh, tmp_path):
    from yolohtli_b200 import _lib
    with pytest.raises(_lib.YolohtliError):
        _lib.load_library(str(tmp_path / "nope.so"))
