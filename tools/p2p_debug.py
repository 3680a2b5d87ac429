import os, sys, ctypes as C
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import yolohtli_b200 as yh
from yolohtli_b200 import host
from yolohtli_b200._lib import lib
from torch.multiprocessing.reductions import reduce_tensor
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
a = torch.full((1024,), float(rank + 1), dtype=torch.float64, device=dev)
flags = torch.zeros(8, dtype=torch.int32, device=dev)
mine = {"a": reduce_tensor(a), "flags": reduce_tensor(flags), "device": local}
torch.cuda.synchronize()
ev = [None] * world
dist.all_gather_object(ev, mine)
nb = 1 - rank
print(rank, "enable peer", lib().yh_enable_peer_access(int(ev[nb]["device"])), flush=True)
fn, args = ev[nb]["a"]; args = list(args); args[6] = local; pa = fn(*args)
fn, args = ev[nb]["flags"]; args = list(args); args[6] = local; pf = fn(*args)
print(rank, "peer tensor device", pa.device, hex(pa.data_ptr()), "mine", hex(a.data_ptr()), flush=True)
torch.cuda.synchronize()
print(rank, "peer read via torch:", pa[:2].cpu().tolist(), flush=True)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
rc = lib().yh_memcpy_async(C.c_void_p(pa.data_ptr() + 512 * 8), C.c_void_p(a.data_ptr()), 256 * 8, st)
torch.cuda.synchronize(); print(rank, "memcpy rc", rc, flush=True)
rc = lib().yh_flag_set(C.c_void_p(pf.data_ptr() + 4), 7, st)
torch.cuda.synchronize(); print(rank, "flag_set rc", rc, flush=True)
rc = lib().yh_flag_wait(C.c_void_p(flags.data_ptr() + 4), 7, C.c_void_p(flags.data_ptr() + 16), st)
torch.cuda.synchronize(); print(rank, "flag_wait rc", rc, flags.tolist(), a[510:516].tolist(), flush=True)
dist.barrier()
del pa, pf
dist.destroy_process_group()
