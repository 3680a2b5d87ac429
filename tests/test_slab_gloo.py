"""CPU, world_size 2 and 3 over gloo: the row-slab partition + halo-exchange logic of
yolohtli_b200.slab (the host side of the multi-GPU path) reproduces the single-domain run bit
for bit.  The CUDA stepper is replaced by a stand-in that runs the plain-C oracle on the
rank's local rows: treating the local array as a whole sheet is wrong only within n rows of a
slab-internal edge after n steps, i.e. inside the ghost rows, so owned rows stay exact."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nx, ny, halo, nsteps, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import oracle_lib
    from yolohtli_b200.slab import SlabRunner
    oracle = oracle_lib.load()
    oracle.set_threads(1)
    pg = oracle.params_default(nx, ny, timeIntOrder=1, lap4=0)

    def stepper(p, n, uA, vA, uB, vB, rows, tb):
        q = p.copy()            # same physics constants; local rows taken as a whole sheet
        q.ny_global, q.jg0 = p.ny, 0
        u, v = oracle.rd_advance(q, n, uA.numpy(), vA.numpy())
        uB.copy_(torch.from_numpy(u.reshape(uB.shape)))
        vB.copy_(torch.from_numpy(v.reshape(vB.shape)))
        return uB, vB

    rng = np.random.default_rng(77)
    u0 = rng.uniform(-0.1, 1.1, (ny, nx))
    v0 = rng.uniform(0.0, 1.0, (ny, nx))
    run = SlabRunner(pg, rank=rank, world=world, halo=halo, device=torch.device("cpu"), stepper=stepper)
    run.load_global(u0, v0)
    run.advance(nsteps)
    u, v = run.owned()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), u=u.numpy(), v=v.numpy(), j0=run.lay.j0, j1=run.lay.j1)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,halo,nsteps", [(2, 4, 12), (3, 2, 7)])
def test_slab_exchange_matches_single_domain(oracle, tmp_path, world, halo, nsteps):
    nx, ny = 40, 67
    mp.spawn(_worker, args=(world, _free_port(), nx, ny, halo, nsteps, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(77)
    u0 = rng.uniform(-0.1, 1.1, (ny, nx))
    v0 = rng.uniform(0.0, 1.0, (ny, nx))
    pg = oracle.params_default(nx, ny, timeIntOrder=1, lap4=0)
    wu, wv = oracle.rd_advance(pg, nsteps, u0, v0)
    rows = 0
    for r in range(world):
        g = np.load(os.path.join(str(tmp_path), f"r{r}.npz"))
        j0, j1 = int(g["j0"]), int(g["j1"])
        assert np.array_equal(g["u"], wu[j0:j1]) and np.array_equal(g["v"], wv[j0:j1]), r
        rows += j1 - j0
    assert rows == ny


def _sr_comm_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import oracle_lib
    from yolohtli_b200.slab import SlabRunner
    oracle = oracle_lib.load()
    nx, ny, H = 16, 30, 7
    pg = oracle.params_default(nx, ny)
    run = SlabRunner(pg, rank=rank, world=world, halo=H, device=torch.device("cpu"), stepper=lambda *a: None)
    l = run.lay
    glob = np.arange(ny * nx, dtype=np.float64).reshape(ny, nx)
    run.u[run.cur][l.own_lo:l.own_hi] = torch.from_numpy(glob[l.j0:l.j1])       # owned rows only: ghosts stale
    run.v[run.cur][l.own_lo:l.own_hi] = torch.from_numpy(-glob[l.j0:l.j1])
    seen = {}

    def fake_steps(nsteps, record):     # the three communication points of SlabRunner._sr_steps
        yield ("exchange", None)
        seen["u"] = run.u[run.cur].clone()
        infos = yield ("gather", torch.tensor([float(rank + 1), 10.0 * rank, 0.5], dtype=torch.float64))
        seen["infos"] = torch.stack(infos)
        rows = torch.zeros(24, dtype=torch.float64)
        rows[rank::world] = 1.0 + rank                                           # one contributor per slot
        yield ("sum", rows)
        seen["rows"] = rows.clone()

    run._sr_steps = fake_steps
    run.advance_sr(1)
    assert torch.equal(seen["u"], torch.from_numpy(glob[l.g0:l.g1])), "ghost rows not refreshed"
    assert seen["infos"].shape == (world, 3)
    assert seen["infos"][:, 0].tolist() == [float(r + 1) for r in range(world)]   # rank order
    want = torch.zeros(24, dtype=torch.float64)
    for r in range(world):
        want[r::world] = 1.0 + r
    assert torch.equal(seen["rows"], want)
    open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    dist.barrier()
    dist.destroy_process_group()


def test_sr_step_collectives_over_gloo(tmp_path):
    """The communication points of the symmetry-reduction slab step (ghost exchange, tip-info
    gather in rank order, row-sum reduction) served by torch.distributed, world_size 2 on CPU."""
    world = 2
    mp.spawn(_sr_comm_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"ok{r}")) for r in range(world))


def test_sr_emulation_harness_serves_the_same_collectives():
    """advance_sr_emulated (the one-process vehicle of tests/test_gpu_sr_slabs.py) answers the three
    communication points exactly like the torch.distributed driver: ghost rows from the neighbours,
    tip infos in rank order, element-wise row sums."""
    from tests import oracle_lib
    from yolohtli_b200.slab import SlabRunner, advance_sr_emulated
    oracle = oracle_lib.load()
    nx, ny, H, world = 16, 45, 7, 3
    pg = oracle.params_default(nx, ny)
    glob = np.arange(ny * nx, dtype=np.float64).reshape(ny, nx)
    runners, seen = [], []
    for rank in range(world):
        run = SlabRunner(pg, rank=rank, world=world, halo=H, device=torch.device("cpu"), stepper=lambda *a: None)
        l = run.lay
        run.u[run.cur][l.own_lo:l.own_hi] = torch.from_numpy(glob[l.j0:l.j1])
        run.v[run.cur][l.own_lo:l.own_hi] = torch.from_numpy(-glob[l.j0:l.j1])
        got = {}

        def fake_steps(nsteps, record, run=run, rank=rank, got=got):
            yield ("exchange", None)
            got["u"], got["v"] = run.u[run.cur].clone(), run.v[run.cur].clone()
            infos = yield ("gather", torch.tensor([float(rank + 1), 10.0 * rank, 0.5], dtype=torch.float64))
            got["infos"] = torch.stack(infos)
            rows = torch.zeros(24, dtype=torch.float64)
            rows[rank::world] = 1.0 + rank
            yield ("sum", rows)
            got["rows"] = rows.clone()

        run._sr_steps = fake_steps
        runners.append(run)
        seen.append(got)
    advance_sr_emulated(runners, 1)
    want_rows = torch.zeros(24, dtype=torch.float64)
    for r in range(world):
        want_rows[r::world] = 1.0 + r
    for run, got in zip(runners, seen):
        l = run.lay
        assert torch.equal(got["u"], torch.from_numpy(glob[l.g0:l.g1])) and torch.equal(got["v"], torch.from_numpy(-glob[l.g0:l.g1]))
        assert got["infos"][:, 0].tolist() == [1.0, 2.0, 3.0]
        assert torch.equal(got["rows"], want_rows)


def test_partition_covers_domain():
    from yolohtli_b200.slab import SlabLayout, partition
    for ny, world in [(16384, 8), (67, 3), (1000, 7)]:
        edges = [partition(ny, world, r) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == ny
        assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
        for r in range(world):
            l = SlabLayout(ny, world, r, 4)
            assert l.g0 == max(0, l.j0 - 4) and l.g1 == min(ny, l.j1 + 4)
            assert (l.up is None) == (r == 0) and (l.down is None) == (r == world - 1)
