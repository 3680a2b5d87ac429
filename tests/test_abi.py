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


REFERENCE_WRAPPERS = [   # mangled names of hostPrototypes.h:22-57 as the reference's translation units export them
    "_Z25reactionDiffusion_wrapperm4dim3S_8stateVarS0_S0_S0_bPbbPdb4int2",
    "_Z11tip_wrapperm4dim3S_8stateVarS0_S0_dibPbPiP7vec5dyn",
    "_Z13slice_wrapperm4dim3S_8stateVar8sliceVarS1_bb6advVariPbPiP7vec5dyni",
    "_Z17Cxy_field_wrapperm4dim3S_6advVar5REAL3S1_Pb",
    "_Z18advFDBFECC_wrapperm4dim3S_8stateVarS0_6advVarS0_S0_S0_Pb",
    "_Z12solve_matrix5REAL3S_Pd",
    "_Z13trapz_wrapper4dim3S_8sliceVarS0_8stateVarPdS2_PiP7vec5dyni",
    "_Z18singleCell_wrapperm4dim3S_8stateVariPdS1_4int2",
    "_Z12sAPD_wrapperm4dim3S_iPdS0_S0_S0_S0_S0_S0_S0_PbS1_b",
    "_Z16countour_wrapperm4dim3S_PdS0_PbS1_PiP6float3fi",
    "_Z7swapSoAP8stateVarS0_",
]


def test_shim_exports_the_full_mangled_set_of_the_reference():
    """Every wrapper symbol of the reference's own translation units (nm of oracle/_ref/libyhref.so when it is
    built here, else the recorded list) is exported by libyolohtli_shim.so under the SAME mangled name, and the
    shim refers to the reference's global `param` weakly (zero-source-change link)."""
    from yolohtli_b200 import _lib
    import subprocess
    from tests import oracle_lib
    shim = subprocess.check_output(["nm", "-D", _lib.SHIM_PATH]).decode()
    exported = {l.split()[-1] for l in shim.splitlines() if " T " in l}
    want = set(REFERENCE_WRAPPERS)
    if os.path.exists(oracle_lib.REF_SO):
        ref = subprocess.check_output(["nm", "-D", "--defined-only", oracle_lib.REF_SO]).decode()
        ref_syms = {l.split()[-1] for l in ref.splitlines()
                    if " T _Z" in l and ("wrapper" in l or "solve_matrix" in l or "swapSoA" in l)}
        assert ref_syms == want, "the recorded list is out of date with the reference build"
    assert want <= exported, want - exported
    assert "_Z16get_rgba_wrapperm4dim3S_iPdPjS1_Pb" in exported          # main.cu:1633 (GL translation unit)
    assert any(l.split()[-2:] == ["w", "param"] for l in shim.splitlines()), "weak reference to `param` missing"


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference headers")
def test_compat_paramvar_layout_matches_the_reference(tmp_path):
    """include/yolohtli_compat.h re-declares paramVar (typeDefinition.cuh:35-125); the shim reads the
    reference's global through it, so size and field offsets must be identical."""
    import subprocess
    body = """
#include <cstdio>
#include <cstddef>
int main() {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(paramVar), offsetof(paramVar, solidSwitch),
         offsetof(paramVar, lap4), offsetof(paramVar, nx), offsetof(paramVar, dt), offsetof(paramVar, invdy),
         offsetof(paramVar, point), offsetof(paramVar, qx4), offsetof(paramVar, tipOffsetX), offsetof(paramVar, tipx),
         offsetof(paramVar, theta));
  return 0;
}
"""
    outs = []
    for name, head in (("ours", f'#include "{ROOT}/include/yolohtli_compat.h"\n'),
                       ("ref", '#include <cuda_runtime.h>\n#include "/root/reference/typeDefinition.cuh"\n')):
        src = tmp_path / f"{name}.cu"
        src.write_text(head.replace("\\n", "\n") + body.replace("\\\\n", "\\n"))
        exe = tmp_path / name
        subprocess.check_call(["nvcc", "-o", str(exe), str(src)], stderr=subprocess.DEVNULL)
        outs.append(subprocess.check_output([str(exe)]).decode().split())
    assert outs[0] == outs[1], outs


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


def test_march_tiling_invariants():
    """Host arithmetic of the marching tile kernel's launcher: strips x bands cover the rows, respect the tile's
    capacity (64 columns, 48 rows, halo = stages on every side) and fill at most one wave where that is possible."""
    import ctypes as C
    from yolohtli_b200 import _lib
    l = _lib.lib()
    for nx, rows, stages, n_sm in [(512, 512, 4, 148), (256, 256, 4, 148), (640, 37, 4, 148), (8, 8, 2, 148), (768, 768, 4, 148),
                                   (500, 131, 2, 148), (1024, 1024, 4, 148), (512, 512, 4, 132), (118, 61, 4, 16)]:
        t = (C.c_int * 2)()
        assert l.yh_rd_tile_march_tiling(nx, rows, stages, n_sm, t) == 0
        sw, bh = t[0], t[1]
        assert sw > 0 and sw % 2 == 0 and sw <= 64 - 2 * stages and 0 < bh <= 48 - 2 * stages, (nx, rows, sw, bh)
        tiles = -(-nx // sw) * -(-rows // bh)
        least = -(-nx // (64 - 2 * stages)) * -(-rows // (48 - 2 * stages))     # tiles at full capacity
        waves, least_waves = -(-tiles // n_sm), -(-least // n_sm)
        assert waves == least_waves, (nx, rows, stages, n_sm, sw, bh, tiles)
    assert l.yh_rd_tile_march_tiling(511, 512, 4, 148, (C.c_int * 2)()) != 0      # odd nx: not served
