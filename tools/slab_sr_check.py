"""Run under torchrun on N GPUs: symmetry-reduction steps on N row slabs (ghost exchange, tip
gather, row-sum all-reduce over NCCL or the NVLink peer transport) equal yh_sim_run_sr on one GPU,
bit for bit -- fields and the (c, phi) history.  YH_TRANSPORT=p2p|nccl picks the ghost transport.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
         --master-port 29517 tools/slab_sr_check.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolohtli_b200 as yh  # noqa: E402
from yolohtli_b200.slab import SlabRunner, partition  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
nx = ny = int(os.environ.get("YH_N", "512"))
nsteps = int(os.environ.get("YH_STEPS", "200"))
# a developed spiral first (standard PDE steps from the cross-field IC, as the reference's users do
# before switching the co-moving frame on); every rank forms the same state on its own GPU
pw = yh.default_params(nx, ny, scale_L=True, timeIntOrder=1, lap4=0)
warm = yh.Sim(pw, device=local)
warm.cross_field_ic()
warm.run(int(os.environ.get("YH_WARM", "12001")), tb_steps=4)
tips = warm.tips()
u0, v0 = warm.get_state()
u0, v0 = u0[0], v0[0]
warm.close()
tx, ty = (float(tips[-1]["x"]), float(tips[-1]["y"])) if len(tips) else (nx / 2.0, ny / 2.0)
p = yh.default_params(nx, ny, reduce_sym=True, scale_L=True, tipOffsetX=160, tipOffsetY=160, tipx0=tx, tipy0=ty)
transport = os.environ.get("YH_TRANSPORT", "nccl")
run = SlabRunner(p, rank=rank, world=world, halo=p.timeIntOrder + 3, device=dev, transport=transport)
run.load_global(u0, v0)
run.sr_setup()
rec = []
run.advance_sr(5, rec)          # warm-up (first-step double solve, first use of every kernel)
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
run.advance_sr(nsteps - 5, rec)
torch.cuda.synchronize()
dist.barrier()
dt = time.perf_counter() - t0
u, v = run.owned()
parts_u = [torch.empty((partition(ny, world, r)[1] - partition(ny, world, r)[0], nx), dtype=torch.float64,
                       device=dev) for r in range(world)]
parts_v = [torch.empty_like(t) for t in parts_u]
dist.all_gather(parts_u, u.contiguous())
dist.all_gather(parts_v, v.contiguous())
if rank == 0:
    sim = yh.Sim(p, device=local)
    sim.set_state(u0[None], v0[None])
    want = sim.run_sr(nsteps)
    gu, gv = sim.get_state()
    sim.close()
    assert np.isfinite(want).all() and np.abs(want[:, :3]).max() > 0, "degenerate run: no drift to compare"
    ok = (np.array_equal(torch.cat(parts_u).cpu().numpy(), gu[0]) and
          np.array_equal(torch.cat(parts_v).cpu().numpy(), gv[0]) and np.array_equal(np.array(rec), want))
    print(f"SR slab check transport={transport} world={world} {nx}x{ny}: {'BITWISE OK' if ok else 'MISMATCH'}; "
          f"{dt / (nsteps - 5) * 1e6:.1f} us per step", flush=True)
    assert ok
assert run.p2p_status() == 0, "p2p flag wait timed out"
run.close()
dist.destroy_process_group()
