"""Run under torchrun on N GPUs: the NCCL row-slab run equals the single-GPU run, bitwise."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolohtli_b200 as yh  # noqa: E402
from yolohtli_b200 import host, synth  # noqa: E402
from yolohtli_b200.slab import SlabRunner  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
nx, ny, nsteps = 2048, 1536, 40
p = yh.default_params(nx, ny, scale_L=True, timeIntOrder=1, lap4=0)
u0, v0 = synth.fibrillation_ic(nx, ny)
transport = os.environ.get("YH_TRANSPORT", "nccl")
run = SlabRunner(p, rank=rank, world=world, halo=4, device=dev, transport=transport)
run.load_global(u0, v0)
run.advance(nsteps, tb=4)
u, v = run.owned()
parts_u = [torch.empty((yh.slab.partition(ny, world, r)[1] - yh.slab.partition(ny, world, r)[0], nx),
                       dtype=torch.float64, device=dev) for r in range(world)]
parts_v = [torch.empty_like(t) for t in parts_u]
dist.all_gather(parts_u, u.contiguous())
dist.all_gather(parts_v, v.contiguous())
if rank == 0:
    gu, gv = torch.cat(parts_u), torch.cat(parts_v)
    uA, vA = torch.as_tensor(u0).to(dev), torch.as_tensor(v0).to(dev)
    uB, vB = torch.zeros_like(uA), torch.zeros_like(vA)
    ru, rv = host.rd_advance(p, nsteps, uA, vA, uB, vB, tb_steps=4)
    torch.cuda.synchronize()
    ok = torch.equal(gu, ru) and torch.equal(gv, rv)
    print(f"slab check transport={transport} world={world}: {'BITWISE OK' if ok else 'MISMATCH'} "
          f"(max |du| = {(gu - ru).abs().max().item():.3e})", flush=True)
    assert ok
assert run.p2p_status() == 0, 'p2p flag wait timed out'
run.close()
dist.destroy_process_group()
