"""Row-slab multi-GPU driver: one process per GPU (torch.distributed for the plumbing), the
nx x ny_global sheet cut into contiguous row slabs, H ghost rows per side refreshed by a
neighbour exchange every H time steps (H = temporal-blocking depth).  The reference is
single-GPU; this is the new multi-GPU layer of section 8e.

Exchange volume per neighbour and direction: H * nx * 8 B * 2 fields every H steps -- at
nx = 16384, H = 4 that is 1 MiB against ~0.5 ms of compute, so the exchange is latency-, not
bandwidth-bound; it is issued on a side stream and overlapped with the interior rows.

The stepping backend is injectable so the partition / exchange logic is testable on CPU with
gloo (tests/test_slab_gloo.py, stand-in stepper); the product backend is the CUDA library.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import host
from ._lib import YhParams


def partition(ny, world, rank):
    """Rows [j0, j1) owned by `rank`: as even as possible, remainder to the low ranks."""
    base, rem = divmod(ny, world)
    j0 = rank * base + min(rank, rem)
    return j0, j0 + base + (1 if rank < rem else 0)


class SlabLayout:
    """Where a rank's rows live.  Local storage holds global rows [g0, g1) = the owned rows
    [j0, j1) plus up to `halo` ghost rows on each side (none beyond the physical boundary,
    where the no-flux mirror applies instead)."""

    def __init__(self, ny_global, world, rank, halo):
        self.ny_global, self.world, self.rank, self.halo = ny_global, world, rank, halo
        self.j0, self.j1 = partition(ny_global, world, rank)
        if world > 1 and (self.j1 - self.j0) < halo:
            raise ValueError("slab thinner than the halo")
        self.g0 = max(0, self.j0 - halo)
        self.g1 = min(ny_global, self.j1 + halo)
        self.ny_local = self.g1 - self.g0
        self.own_lo = self.j0 - self.g0          # local row of the first owned row
        self.own_hi = self.j1 - self.g0
        self.up = rank - 1 if rank > 0 else None             # neighbour holding rows below j0
        self.down = rank + 1 if rank < world - 1 else None   # neighbour holding rows from j1

    def local_params(self, p_global):
        p = YhParams()
        C.memmove(C.byref(p), C.byref(p_global), C.sizeof(YhParams))
        p.ny = self.ny_local
        p.ny_global = self.ny_global
        p.jg0 = self.g0
        return p


class SlabRunner:
    def __init__(self, p_global, rank=None, world=None, halo=4, device=None, stepper=None, transport="nccl"):
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.lay = SlabLayout(p_global.ny, self.world, self.rank, halo)
        self.p = self.lay.local_params(p_global)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.halo = halo
        # rows consumed per time step: one per stage, two with the anisotropic no-flux corner terms
        # (reactionDiffusion.cu:290-304 read rows j +- 2 at the x edges)
        rad = 2 if (p_global.anisotropy and p_global.neumannBC and not p_global.solidSwitch) else 1
        self.K = max(1, p_global.timeIntOrder) * rad
        self.nx = p_global.nx
        shape = (self.lay.ny_local, self.nx)
        self.u = [torch.zeros(shape, dtype=torch.float64, device=self.device) for _ in range(2)]
        self.v = [torch.zeros(shape, dtype=torch.float64, device=self.device) for _ in range(2)]
        self.cur = 0
        self.solid = None            # obstacle mask rows of this slab (ghosts included) ...
        self.solid_flags = 0         # ... or their precomputed patterns (masked Euler)
        self.stepper = stepper or self._cuda_stepper
        self.count = 0
        self.comm_stream = (torch.cuda.Stream(device=self.device, priority=-1)
                            if self.device.type == "cuda" else None)
        self.transport = transport if (self.world > 1 and self.device.type == "cuda") else "nccl"
        self.seq = 0
        if self.transport == "p2p":
            self._setup_p2p()

    # -- state ---------------------------------------------------------------------------
    def load_global(self, u_glob, v_glob):
        """Every rank slices its rows (ghosts included) out of the same global arrays."""
        l = self.lay
        self.u[self.cur].copy_(torch.as_tensor(u_glob[l.g0:l.g1]).to(self.device))
        self.v[self.cur].copy_(torch.as_tensor(v_glob[l.g0:l.g1]).to(self.device))
        self.count = 0   # fresh host data: the next pass treats it as raw

    def set_solid(self, mask_glob):
        """Obstacle mask of the whole domain (1 = tissue, main.cu:676-680); every rank keeps its
        rows.  For the masked Euler mode the neighbourhood patterns are derived once."""
        l = self.lay
        m = torch.as_tensor(mask_glob[l.g0:l.g1].astype('uint8')).to(self.device).contiguous()
        self.solid, self.solid_flags = m, 0
        if self.device.type == "cuda" and self.p.solidSwitch and self.p.timeIntOrder == 1 and self.p.neumannBC:
            pat = torch.empty_like(m)
            host.rd_mask_patterns(self.p, m, pat)
            self.solid, self.solid_flags = pat, host.RD_SOLID_IS_PATTERNS

    def owned(self):
        l = self.lay
        return self.u[self.cur][l.own_lo:l.own_hi], self.v[self.cur][l.own_lo:l.own_hi]

    # -- halo exchange -------------------------------------------------------------------
    def _ops(self):
        l, H = self.lay, self.halo
        u, v = self.u[self.cur], self.v[self.cur]
        ops = []
        if l.up is not None:      # my first H owned rows -> up's lower ghosts; its last H -> mine
            ops += [dist.P2POp(dist.isend, u[l.own_lo:l.own_lo + H], l.up),
                    dist.P2POp(dist.isend, v[l.own_lo:l.own_lo + H], l.up),
                    dist.P2POp(dist.irecv, u[l.own_lo - H:l.own_lo], l.up),
                    dist.P2POp(dist.irecv, v[l.own_lo - H:l.own_lo], l.up)]
        if l.down is not None:
            ops += [dist.P2POp(dist.isend, u[l.own_hi - H:l.own_hi], l.down),
                    dist.P2POp(dist.isend, v[l.own_hi - H:l.own_hi], l.down),
                    dist.P2POp(dist.irecv, u[l.own_hi:l.own_hi + H], l.down),
                    dist.P2POp(dist.irecv, v[l.own_hi:l.own_hi + H], l.down)]
        return ops

    # -- NVLink peer-to-peer transport ------------------------------------------------------
    # Every rank exports its four state arrays and a flag word through CUDA IPC (torch's own
    # reduce_tensor machinery); a rank WRITES its fresh edge rows straight into the neighbours'
    # ghost rows (cudaMemcpyAsync through the peer mapping = NVLink stores), then releases a
    # sequence number into the neighbour's flag; the neighbour's stream acquires it
    # (yh_flag_wait) before reading the ghosts.  Write-after-read safety follows from the same
    # chain: a rank cannot reach exchange k before it consumed its neighbour's exchange k-1,
    # which the neighbour issued after it had finished reading the ghosts that k overwrites.
    def _setup_p2p(self):
        from torch.multiprocessing.reductions import reduce_tensor
        self.flags = torch.zeros(8, dtype=torch.int32, device=self.device)   # [0] from up, [1] from down, [4] status
        mine = {}
        for name, t in (("u0", self.u[0]), ("u1", self.u[1]), ("v0", self.v[0]), ("v1", self.v[1]),
                        ("flags", self.flags)):
            fn, args = reduce_tensor(t)
            mine[name] = (fn, args)
        mine["device"] = self.device.index
        torch.cuda.synchronize(self.device)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine)
        self.peer = {}
        from ._lib import lib
        for nb in (self.lay.up, self.lay.down):
            if nb is None:
                continue
            # the IPC mapping is opened in a context of the exporter's device; kernels and copies of
            # THIS device reach it over NVLink once peer access is enabled
            with torch.cuda.device(self.device):
                host.check(lib().yh_enable_peer_access(int(everyone[nb]["device"])))
            d = {}
            for name, val in everyone[nb].items():
                if name == "device":
                    continue
                fn, args = val
                args = list(args)
                args[6] = self.device.index  # open the handle with THIS device current: the mapping is
                d[name] = fn(*args)          # then peer-accessible from our kernels (lazy peer access)
            self.peer[nb] = d
        self._peer_lay = {nb: SlabLayout(self.lay.ny_global, self.world, nb, self.halo) for nb in self.peer}
        dist.barrier()

    def _exchange_p2p(self):
        from ._lib import lib
        l, H = self.lay, self.halo
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        self.seq += 1
        row_bytes = self.nx * 8
        cur = self.cur
        for nb, my_rows, flag_idx_at_peer in ((l.up, (l.own_lo, l.own_lo + H), 1), (l.down, (l.own_hi - H, l.own_hi), 0)):
            if nb is None:
                continue
            pl = self._peer_lay[nb]
            # my first H owned rows are the peer's lower ghosts (rows own_hi..own_hi+H there); my last H
            # owned rows are the peer's upper ghosts (rows own_lo-H..own_lo there)
            dst_row = pl.own_hi if nb == l.up else pl.own_lo - H
            for f in ("u", "v"):
                src = (self.u if f == "u" else self.v)[cur]
                dst = self.peer[nb][f + str(cur)]
                host.check(lib().yh_memcpy_async(C.c_void_p(dst.data_ptr() + dst_row * row_bytes),
                                                 C.c_void_p(src.data_ptr() + my_rows[0] * row_bytes),
                                                 H * row_bytes, st))
            host.check(lib().yh_flag_set(C.c_void_p(self.peer[nb]["flags"].data_ptr() + 4 * flag_idx_at_peer),
                                         self.seq, st))
        for nb, idx in ((l.up, 0), (l.down, 1)):
            if nb is not None:
                host.check(lib().yh_flag_wait(C.c_void_p(self.flags.data_ptr() + 4 * idx), self.seq,
                                              C.c_void_p(self.flags.data_ptr() + 16), st))

    def p2p_status(self):
        return int(self.flags[4].item()) if self.transport == "p2p" else 0

    def close(self):
        if self.transport == "p2p" and getattr(self, "peer", None):
            import gc
            torch.cuda.synchronize(self.device)
            dist.barrier()
            self.peer = {}            # drop the IPC mappings before the exporters exit
            gc.collect()
            torch.cuda.ipc_collect()
            dist.barrier()

    def exchange(self):
        if self.transport == "p2p":
            return self._exchange_p2p()
        ops = self._ops()
        if not ops:
            return
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    # -- stepping ------------------------------------------------------------------------
    def _cuda_stepper(self, p, nsteps, uA, vA, uB, vB, rows, tb):
        # after the first pass every value was written by the library: no -0.0 can be present
        flags = (host.RD_INPUT_CANONICAL if self.count > 0 else 0) | self.solid_flags
        return host.rd_advance(p, nsteps, uA, vA, uB, vB, tb_steps=tb, rows=rows, flags=flags, solid=self.solid)

    def advance(self, nsteps, tb=0, overlap=None):
        """nsteps time steps: ghosts refreshed, then up to `halo` steps per exchange.
        overlap (default: on for CUDA runs with neighbours): the edge rows of a slab are stepped
        first on a high-priority stream and sent while the interior rows are still computing."""
        if overlap is None:
            overlap = self.world > 1 and self.comm_stream is not None and self.stepper == self._cuda_stepper
        if overlap and self.K == 1:
            return self._advance_overlapped(nsteps, tb)
        l = self.lay
        left = nsteps
        while left > 0:
            n = min(max(1, self.halo // self.K), left)
            if self.world > 1:
                self.exchange()
            c, o = self.cur, self.cur ^ 1
            ru, rv = self.stepper(self.p, n, self.u[c], self.v[c], self.u[o], self.v[o],
                                  (l.own_lo, l.own_hi), tb)
            self.cur = o if ru is self.u[o] else c
            left -= n
            self.count += n

    # -- symmetry-reduction (co-moving frame) steps on row slabs: SURVEY 8e, "SR mode" ------------
    # One step (main.cu:894-954; yh_sim_run_sr on a whole sheet) needs timeIntOrder + 3 ghost rows:
    #   exchange ghosts of (u, v)^n
    #   RD            rows own-3 .. own+3 -> (u*, v*), velTan                       [yh_rd_step]
    #   tips          owned rows, u^n vs u*; every rank learns the LAST tip of the    [yh_tip_track_rows]
    #                 concatenated list (slabs are contiguous in j, so rank order = cell order)
    #   integrals     row sums of the owned disc rows; element-wise sum over ranks    [yh_sr_integral_rows]
    #                 (one non-zero contributor per row: exact), canonical closing    [yh_sr_integrals_close]
    #   solve         3x3 on the host of EVERY rank (same 12 numbers -> same c)       [yh_solve_matrix]
    #   BFECC         owned rows of u* -> (u, v)^{n+1} in the frame moving with c     [yh_advect_bfecc_cphi_rows]
    # Written as a generator that yields at the three communication points so that the same text
    # runs over torch.distributed (advance_sr) and, for tests, over N runners emulated in one
    # process (advance_sr_emulated).  N slabs == the whole-sheet yh_sim_run_sr bit for bit.
    def sr_setup(self, tip_capacity=65536):
        need = self.K + 3
        if self.world > 1 and self.halo < need:
            raise ValueError(f"symmetry-reduction steps need {need} ghost rows (timeIntOrder + 3), have {self.halo}")
        shape = (self.lay.ny_local, self.nx)
        z = lambda: torch.zeros(shape, dtype=torch.float64, device=self.device)   # noqa: E731
        self.vt = [z(), z()]            # velTan and the advection field start at zero (main.cu:431-434)
        self.adv = [z(), z()]
        self.tip_capacity = tip_capacity
        self.tip_count = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.tip_vec = torch.zeros(tip_capacity * 20, dtype=torch.uint8, device=self.device)
        self.sr_rows = torch.zeros(12 * host.sr_disc_slots(self.p), dtype=torch.float64, device=self.device)
        self.c = [0.0, 0.0, 0.0]
        self.phi = [0.0, 0.0, 0.0]
        self.sr_count = 0
        self.sr_tips = []               # this rank's tips of the last step (numpy records)

    def _sr_local_tip_info(self):
        """[n, x_last, y_last] of this rank's list (float32 values carried exactly in float64)."""
        n = int(self.tip_count.item())
        if n > self.tip_capacity:
            raise RuntimeError("tip list overflow on a slab")
        info = torch.zeros(3, dtype=torch.float64)
        info[0] = n
        self.sr_tips = host.tips_to_numpy(self.tip_count, self.tip_vec)
        if n > 0:
            info[1], info[2] = float(self.sr_tips[-1]["x"]), float(self.sr_tips[-1]["y"])
        return info

    def _sr_steps(self, nsteps, record):
        l, p = self.lay, self.p
        own = (l.own_lo, l.own_hi)
        for it in range(nsteps):
            if self.world > 1:
                yield ("exchange", None)
            c, o = self.cur, self.cur ^ 1
            rows_rd = (max(0, l.own_lo - 3), min(l.ny_local, l.own_hi + 3))
            host.rd_step(p, self.u[c], self.v[c], self.u[o], self.v[o], velTan=(self.vt[0], self.vt[1]),
                         solid=self.solid, rows=rows_rd)
            host.tip_track_rows(p, self.u[o], self.u[c], self.tip_count, self.tip_vec, own,
                                t=p.dt * self.sr_count, capacity=self.tip_capacity)
            infos = yield ("gather", self._sr_local_tip_info())
            centre = (p.tipx0, p.tipy0)        # count == 0, or no tip anywhere (defect B3: keep the centre)
            if self.sr_count != 0:
                for info in infos:             # rank order = ascending j: the last non-empty list wins
                    if info[0] > 0:
                        centre = (float(info[1]), float(info[2]))
            if record is not None:
                record.append(list(self.c) + list(self.phi))          # main.cu:902-903
            passes = 2 if self.sr_count == 0 else 1                    # first step: main.cu:910-921
            for k in range(passes):
                host.sr_integral_rows(p, self.u[c], self.v[c], self.vt[0], self.vt[1], self.adv[0], self.adv[1],
                                      centre, own, self.sr_rows)
                yield ("sum", self.sr_rows)
                I = host.sr_integrals_close(p, self.sr_rows)
                self.c = list(host.solve_matrix(self.c, self.phi, I))
                if passes == 2 and k == 0:
                    host.cxy_field(p, self.adv[0], self.adv[1], self.c, self.phi, solid=self.solid)
            host.advect_bfecc_cphi_rows(p, self.u[o], self.v[o], self.u[c], self.v[c], self.c, self.phi, own,
                                        adv_x=self.adv[0], adv_y=self.adv[1], solid=self.solid)
            self.phi = [self.phi[q] + self.c[q] * p.dt for q in range(3)]   # main.cu:936-938
            self.sr_count += 1
            self.count += 1

    def advance_sr(self, nsteps, record=None):
        """nsteps symmetry-reduction steps over torch.distributed; record (a list) receives
        [c, phi] per step as pushed to clist / philist."""
        gen = self._sr_steps(nsteps, record)
        reply = None
        while True:
            try:
                what, arg = gen.send(reply)
            except StopIteration:
                return
            reply = None
            if what == "exchange":
                self.exchange()
            elif what == "gather":
                t = arg.to(self.device) if self.device.type == "cuda" else arg
                out = [torch.zeros_like(t) for _ in range(self.world)]
                if self.world > 1:
                    dist.all_gather(out, t)
                else:
                    out = [t]
                reply = [x.cpu() for x in out]
            elif what == "sum" and self.world > 1:
                dist.all_reduce(arg, op=dist.ReduceOp.SUM)

    # -- overlapped schedule ----------------------------------------------------------------
    #   edge stream (high priority):  wait ghosts(k), interior(k-1) -> step edge rows -> NCCL
    #                                 send/recv of the fresh edge rows = ghosts(k+1)
    #   main stream:                  wait edges(k-1)               -> step interior rows
    # Interior rows are >= halo rows away from the slab edges, so they never read a ghost row
    # and run concurrently with the exchange.  One HBM pass (n = halo steps) per block.
    def _advance_overlapped(self, nsteps, tb):
        l, H = self.lay, self.halo
        B = max(H, min(128, (l.own_hi - l.own_lo) // 4))     # edge band height
        main = torch.cuda.current_stream(self.device)
        edge = self.comm_stream
        top = (l.own_lo, l.own_lo + B) if l.up is not None else None
        bot = (l.own_hi - B, l.own_hi) if l.down is not None else None
        inner = (top[1] if top else l.own_lo, bot[0] if bot else l.own_hi)
        ev_int, ev_edge = torch.cuda.Event(), torch.cuda.Event()
        ev_int.record(main)
        ev_edge.record(main)
        left = nsteps
        first = True
        while left > 0:
            n = max(t for t in (1, 2, 4) if t <= min(H, left, tb or 4))   # exactly one HBM pass
            c, o = self.cur, self.cur ^ 1
            flags = (host.RD_INPUT_CANONICAL if self.count > 0 else 0) | self.solid_flags
            with torch.cuda.stream(edge):
                edge.wait_event(ev_int)                      # interior(k-1) wrote rows the edges read
                if first:                                    # ghosts of the initial state
                    self.exchange()
                    first = False
                for rows in (top, bot):
                    if rows is not None:
                        host.rd_advance(self.p, n, self.u[c], self.v[c], self.u[o], self.v[o], tb_steps=n,
                                        rows=rows, flags=flags, solid=self.solid)
                ev_edge_next = torch.cuda.Event()
                ev_edge_next.record(edge)
            main.wait_event(ev_edge)                         # edges(k-1) wrote rows the interior reads
            if inner[1] > inner[0]:
                host.rd_advance(self.p, n, self.u[c], self.v[c], self.u[o], self.v[o], tb_steps=n,
                                rows=inner, flags=flags, solid=self.solid)
            ev_int = torch.cuda.Event()
            ev_int.record(main)
            ev_edge = ev_edge_next
            self.cur = o
            left -= n
            self.count += n
            if left > 0:
                with torch.cuda.stream(edge):
                    self.exchange()                          # fresh edge rows -> neighbours' ghosts
        main.wait_event(ev_edge)
        main.wait_stream(edge)


def advance_sr_emulated(runners, nsteps, records=None):
    """The symmetry-reduction steps of N SlabRunners (world = N, ranks 0..N-1, one process, one
    device): the same per-rank text as SlabRunner.advance_sr with the three collectives served
    locally -- ghost rows copied between the runners' buffers, tip infos gathered in rank order,
    row sums added element-wise.  Test vehicle for the multi-GPU logic on a single GPU."""
    n = len(runners)
    gens = [r._sr_steps(nsteps, None if records is None else records[i]) for i, r in enumerate(runners)]
    replies = [None] * n
    while True:
        reqs = []
        for i, g in enumerate(gens):
            try:
                reqs.append(g.send(replies[i]))
            except StopIteration:
                reqs.append(None)
        if all(q is None for q in reqs):
            return
        assert all(q is not None for q in reqs) and len({q[0] for q in reqs}) == 1, "ranks out of step"
        what = reqs[0][0]
        replies = [None] * n
        if what == "exchange":
            for i, r in enumerate(runners):
                l, H = r.lay, r.halo
                if l.down is not None:
                    d = runners[l.down]
                    for a, b in ((r.u[r.cur], d.u[d.cur]), (r.v[r.cur], d.v[d.cur])):
                        b[d.lay.own_lo - H:d.lay.own_lo] = a[l.own_hi - H:l.own_hi]
                        a[l.own_hi:l.own_hi + H] = b[d.lay.own_lo:d.lay.own_lo + H]
        elif what == "gather":
            infos = [q[1].clone() for q in reqs]
            replies = [infos] * n
        elif what == "sum":
            total = reqs[0][1].clone()
            for q in reqs[1:]:
                total += q[1]
            for q in reqs:
                q[1].copy_(total)


# ---------------------------------------------------------------------------------------------------
# ctypes mirrors of the C++ multi-GPU driver (include/yolohtli_slab.h, csrc/slab.cu).  The partition,
# the overlap schedule, the NVLink halo exchange and the graph replay all live in the library; Python
# only moves the 256-byte handles between processes (any transport) and owns the host buffers.
# ---------------------------------------------------------------------------------------------------
def _hostptr(a):
    """numpy array or (pinned) torch CPU tensor -> void*"""
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous()
        return C.c_void_p(a.data_ptr())
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class Slab:
    """One row slab of an nx x ny sheet on one GPU (yh_slab)."""

    def __init__(self, p_global, rank, world, halo=4, device=0):
        from ._lib import lib
        self._lib = lib()
        self._h = C.c_void_p()
        host.check(self._lib.yh_slab_create(C.byref(self._h), C.byref(p_global), rank, world, halo, device))
        self.rank, self.world, self.nx = rank, world, p_global.nx
        j0, j1, g0, g1 = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        host.check(self._lib.yh_slab_layout(self._h, C.byref(j0), C.byref(j1), C.byref(g0), C.byref(g1)))
        self.j0, self.j1, self.g0, self.g1 = j0.value, j1.value, g0.value, g1.value

    def export(self):
        from ._lib import SLAB_HANDLE_BYTES
        buf = C.create_string_buffer(SLAB_HANDLE_BYTES)
        host.check(self._lib.yh_slab_export(self._h, buf))
        return buf.raw

    def connect(self, handle_up, handle_down):
        host.check(self._lib.yh_slab_connect(self._h, handle_up, handle_down))

    def connect_over(self, all_gather_bytes):
        """all_gather_bytes(my_bytes) -> list of every rank's bytes (torch.distributed, MPI, ...)."""
        everyone = all_gather_bytes(self.export())
        self.connect(everyone[self.rank - 1] if self.rank > 0 else None,
                     everyone[self.rank + 1] if self.rank < self.world - 1 else None)

    def set_state(self, u_rows, v_rows, with_ghosts=False):
        host.check(self._lib.yh_slab_set_state(self._h, _hostptr(u_rows), _hostptr(v_rows), int(with_ghosts)))

    def get_state(self, out=None):
        import numpy as np
        if out is None:
            out = (np.empty((self.j1 - self.j0, self.nx)), np.empty((self.j1 - self.j0, self.nx)))
        host.check(self._lib.yh_slab_get_state(self._h, _hostptr(out[0]), _hostptr(out[1])))
        return out

    def set_solid(self, mask_rows):
        import numpy as np
        m = np.ascontiguousarray(mask_rows, dtype=np.uint8)
        assert m.shape == (self.g1 - self.g0, self.nx)
        host.check(self._lib.yh_slab_set_solid(self._h, m.ctypes.data_as(C.c_void_p)))

    def advance(self, nsteps, tb=0):
        host.check(self._lib.yh_slab_advance(self._h, nsteps, tb))

    def sync(self):
        host.check(self._lib.yh_slab_sync(self._h))

    def stream(self):
        """The slab's main stream as a torch ExternalStream (for CUDA-event timing)."""
        import torch
        return torch.cuda.ExternalStream(int(self._lib.yh_slab_stream(self._h)))

    def checksum(self):
        a, b = C.c_ulonglong(), C.c_ulonglong()
        host.check(self._lib.yh_slab_checksum(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def run_host(self, u_in, v_in, u_out, v_out, nsteps, tb=0):
        host.check(self._lib.yh_slab_run_host(self._h, _hostptr(u_in), _hostptr(v_in), _hostptr(u_out),
                                              _hostptr(v_out), nsteps, tb))

    def close(self):
        if self._h:
            self._lib.yh_slab_destroy(self._h)
            self._h = C.c_void_p()


class SlabGroup:
    """N slabs driven by one process (yh_slab_group): devices[r] holds slab r."""

    def __init__(self, p_global, devices, halo=4):
        from ._lib import lib
        self._lib = lib()
        self._h = C.c_void_p()
        self.n = len(devices)
        self.shape = (p_global.ny, p_global.nx)
        arr = (C.c_int * self.n)(*devices)
        host.check(self._lib.yh_slab_group_create(C.byref(self._h), C.byref(p_global), self.n, arr, halo))

    def set_state(self, u, v):
        host.check(self._lib.yh_slab_group_set_state(self._h, _hostptr(u), _hostptr(v)))

    def set_solid(self, mask):
        import numpy as np
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        host.check(self._lib.yh_slab_group_set_solid(self._h, m.ctypes.data_as(C.c_void_p)))

    def advance(self, nsteps, tb=0):
        host.check(self._lib.yh_slab_group_advance(self._h, nsteps, tb))

    def advance_sr(self, nsteps):
        """Symmetry-reduction steps (yh_slab_group_advance_sr); returns the (c, phi) record, nsteps x 6."""
        import numpy as np
        rec = np.zeros((nsteps, 6))
        host.check(self._lib.yh_slab_group_advance_sr(self._h, nsteps, _hostptr(rec)))
        return rec

    def get_state(self):
        import numpy as np
        u, v = np.empty(self.shape), np.empty(self.shape)
        host.check(self._lib.yh_slab_group_get_state(self._h, _hostptr(u), _hostptr(v)))
        return u, v

    def checksum(self):
        su = sv = 0
        for r in range(self.n):
            a, b = C.c_ulonglong(), C.c_ulonglong()
            host.check(self._lib.yh_slab_checksum(C.c_void_p(self._lib.yh_slab_group_member(self._h, r)),
                                                  C.byref(a), C.byref(b)))
            su, sv = (su + a.value) % (1 << 64), (sv + b.value) % (1 << 64)
        return su, sv

    def close(self):
        if self._h:
            self._lib.yh_slab_group_destroy(self._h)
            self._h = C.c_void_p()
