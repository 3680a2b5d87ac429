"""GPU: the C++ multi-GPU row-slab driver (include/yolohtli_slab.h, csrc/slab.cu).  N slabs == one
sheet == the plain-C oracle, bit for bit -- driven from a C++ host with no Python in the loop
(tests/slab_driver.cu), through the ctypes mirror, and across PROCESSES over CUDA IPC handles.  On a
one-GPU box every slab lives on the same device: the peer mappings, flags, streams and graphs are the
ones the NVLink path uses."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from yolohtli_b200 import synth  # noqa: E402
from yolohtli_b200.slab import SlabGroup  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "yolohtli_b200", "lib", "yh_slab_driver")


@pytest.mark.parametrize("args", ["512 512 2 203 euler", "768 640 3 131 euler", "512 512 2 37 rk4lap4",
                                  "640 512 3 150 eulerholes", "2048 2048 2 403 euler", "512 520 8 77 euler"])
def test_cpp_host_drives_slabs_bitwise(args):
    """tests/slab_driver.cu: a C++ main() over the C ABI -- N slabs vs the single-device driver."""
    assert os.path.exists(DRIVER), "build the library first (make -C yolohtli_b200/csrc)"
    ndev = str(min(torch.cuda.device_count(), int(args.split()[2])))
    r = subprocess.run([DRIVER] + args.split() + [ndev], capture_output=True, text=True, timeout=300)
    print(r.stdout.strip())
    assert r.returncode == 0 and "slab_driver PASS" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("args", ["256 1024 1 203 euler", "256 1024 2 203 euler", "256 1200 3 171 euler", "128 1024 4 150 euler",
                                  "256 1024 2 61 rk4lap4", "256 1100 3 47 rk4lap4", "2048 4096 2 403 euler"])
def test_cpp_host_pipelined_run_host_bitwise(args):
    """yh_slab_group_run_host with the copies hidden behind the time steps (skewed chunks, edge wedges caught up
    with one exchange per level): host buffers in -> steps -> host buffers out, bit-identical to one sheet."""
    assert os.path.exists(DRIVER), "build the library first (make -C yolohtli_b200/csrc)"
    ndev = str(min(torch.cuda.device_count(), int(args.split()[2])))
    r = subprocess.run([DRIVER] + args.split() + [ndev, "pipe"], capture_output=True, text=True, timeout=300)
    print(r.stdout.strip())
    assert r.returncode == 0 and "slab_driver PASS" in r.stdout and "levels per chunk" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("args", ["256 256 1 40 sr", "256 256 2 60 sr", "256 384 3 40 sr", "512 512 2 100 sr", "1024 1024 2 30 sr"])
def test_cpp_host_symmetry_reduction_on_slabs(args):
    """yh_slab_group_advance_sr: display()'s symmetry-reduction branch on row slabs from a C++ host -- fields and the
    (c, phi) record bit for bit those of yh_sim_run_sr on one sheet, from a spiral the program grows itself."""
    assert os.path.exists(DRIVER), "build the library first (make -C yolohtli_b200/csrc)"
    ndev = str(min(torch.cuda.device_count(), int(args.split()[2])))
    r = subprocess.run([DRIVER] + args.split() + [ndev], capture_output=True, text=True, timeout=300)
    print(r.stdout.strip())
    assert r.returncode == 0 and "slab_driver PASS" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("mode,n,world,steps", [("euler", 384, 2, 100), ("euler", 384, 3, 64), ("rk4lap4", 256, 2, 25),
                                               ("rk2", 256, 3, 12), ("euler_holes", 512, 2, 80)])
def test_slab_group_vs_oracle(oracle, yh, mode, n, world, steps):
    kw = {"euler": dict(timeIntOrder=1, lap4=0), "rk4lap4": {}, "rk2": dict(timeIntOrder=2),
          "euler_holes": dict(timeIntOrder=1, lap4=0, solidSwitch=1)}[mode]
    p = yh.default_params(n, n, **kw)
    u0, v0 = synth.cross_field_ic(n, n)
    mask = None
    if mode == "euler_holes":
        mask = synth.hole_mask(n, seed=3)
        u0, v0 = u0 * mask, v0 * mask
    g = SlabGroup(p, [r % torch.cuda.device_count() for r in range(world)], halo=4)
    if mask is not None:
        g.set_solid(mask)
    g.set_state(u0, v0)
    g.advance(steps)
    gu, gv = g.get_state()
    su, sv = g.checksum()
    g.close()
    wu, wv = oracle.rd_advance(p, steps, u0, v0, solid=mask)
    assert np.array_equal(gu, wu) and np.array_equal(gv, wv)
    assert su == int(wu.view(np.uint64).sum(dtype=np.uint64)) and sv == int(wv.view(np.uint64).sum(dtype=np.uint64))


def test_slabs_across_processes_over_ipc_handles():
    """One process per slab (torchrun, gloo for the 256-byte handles), all on the visible GPUs: the
    export / connect path the one-rank-per-GPU bench uses.  Rank 0 compares with the single-device run."""
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "slab_ipc_check.py"),
           "512", "230"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    print(r.stdout[-2000:])
    assert r.returncode == 0 and "slab_ipc_check PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
