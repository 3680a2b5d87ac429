"""One process per slab: N row slabs of the C++ driver (include/yolohtli_slab.h) wired over CUDA IPC
handles == the single-device run, bit for bit, plus the order-independent global checksum.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         tools/slab_ipc_check.py [n=512] [steps=230] [mode=euler|rk4lap4]

Ranks share the visible GPUs round-robin (N ranks on one GPU is fine: IPC works within a device)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolohtli_b200 as yh  # noqa: E402
from yolohtli_b200 import synth  # noqa: E402
from yolohtli_b200.slab import Slab  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 230
mode = sys.argv[3] if len(sys.argv) > 3 else "euler"
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = rank % torch.cuda.device_count()
torch.cuda.set_device(dev)
dist.init_process_group("gloo")
yh.load_library()
p = yh.default_params(n, n, scale_L=True, **(dict(timeIntOrder=1, lap4=0) if mode == "euler" else {}))
u0, v0 = synth.fibrillation_ic(n, n) if n >= 1024 else synth.cross_field_ic(n, n)


def gather(b):
    out = [None] * world
    dist.all_gather_object(out, b)
    return out


s = Slab(p, rank, world, halo=4, device=dev)
s.connect_over(gather)
mine_u, mine_v = np.ascontiguousarray(u0[s.j0:s.j1]), np.ascontiguousarray(v0[s.j0:s.j1])
out_u, out_v = np.empty_like(mine_u), np.empty_like(mine_v)
dist.barrier()
s.run_host(mine_u, mine_v, out_u, out_v, steps // 2)          # owned rows only: ghosts come from the neighbours
s.advance(steps - steps // 2)                                  # continue from device-resident state
s.get_state(out=(out_u, out_v))
cs = s.checksum()
parts = gather((out_u, out_v, cs))
dist.barrier()
s.close()
if rank == 0:
    sim = yh.Sim(p)
    sim.set_state(u0[None], v0[None])
    sim.run(steps, tb_steps=4)
    wu, wv = (a[0] for a in sim.get_state())
    sim.close()
    gu = np.concatenate([q[0] for q in parts]); gv = np.concatenate([q[1] for q in parts])
    su = sum(q[2][0] for q in parts) % (1 << 64); sv = sum(q[2][1] for q in parts) % (1 << 64)
    ok = np.array_equal(gu, wu) and np.array_equal(gv, wv) and su == int(wu.view(np.uint64).sum(dtype=np.uint64)) \
        and sv == int(wv.view(np.uint64).sum(dtype=np.uint64))
    print(f"slab_ipc_check {'PASS' if ok else 'FAIL'} {n}x{n} {mode} world={world} steps={steps} checksum_u={su:016x}", flush=True)
dist.destroy_process_group()
