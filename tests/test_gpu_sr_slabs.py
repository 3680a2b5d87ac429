"""Symmetry-reduction mode on row slabs ("SR mode"): the slab forms of the tip
tracker, the phase-condition integrals, the frame-velocity field and the BFECC advection
reproduce their whole-sheet forms bit for bit, and N slabs stepped with ghost-row exchange,
tip gather and row-sum reduction (emulated on one GPU, same per-rank text as the
torch.distributed driver) == yh_sim_run_sr on the whole sheet: fields, (c, phi) history, tips."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from yolohtli_b200 import host  # noqa: E402
from yolohtli_b200.slab import SlabLayout, SlabRunner, advance_sr_emulated  # noqa: E402

from .test_gpu_aux import dev, fields, wavy  # noqa: E402


def zeros_like_rows(ny, nx):
    return torch.zeros((ny, nx), dtype=torch.float64, device="cuda")


@pytest.mark.parametrize("world", [2, 3])
def test_tip_rows_concatenate_to_the_whole_list(oracle, world):
    nx, ny = 333, 257
    p = oracle.params_default(nx, ny)
    past, present = wavy(nx, ny), wavy(nx, ny, 0.29, 0.31, 0.4)
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    vec = torch.zeros(host.TIPVECSIZE * 20, dtype=torch.uint8, device="cuda")
    host.tip_track(p, dev(past), dev(present), cnt, vec, t=1.5)
    whole = host.tips_to_numpy(cnt, vec)
    assert len(whole) > 20
    parts = []
    for r in range(world):
        l = SlabLayout(ny, world, r, 1)          # one ghost row: a cell reads the row above it
        q = l.local_params(p)
        host.tip_track_rows(q, dev(past[l.g0:l.g1]), dev(present[l.g0:l.g1]), cnt, vec, (l.own_lo, l.own_hi), t=1.5)
        parts.append(host.tips_to_numpy(cnt, vec))
    got = np.concatenate(parts)
    assert got.tobytes() == whole.tobytes()
    assert min(len(x) for x in parts) > 0, "every slab should contribute"


@pytest.mark.parametrize("world", [2, 3])
def test_sr_pieces_on_row_slabs(oracle, world):
    nx = ny = 128
    p = oracle.params_default(nx, ny, reduce_sym=True, tipOffsetX=40, tipOffsetY=40, tipx0=60.0, tipy0=66.0)
    u, v = fields(nx, ny, 21)
    vtu, vtv = fields(nx, ny, 22)
    c, phi = [0.13, -0.21, 0.04], [0.3, -0.1, 0.77]
    du, dv, dvtu, dvtv = dev(u), dev(v), dev(vtu), dev(vtv)
    # whole sheet
    ax, ay = zeros_like_rows(ny, nx), zeros_like_rows(ny, nx)
    host.cxy_field(p, ax, ay, c, phi)
    I_whole = host.sr_integrals(p, du, dv, dvtu, dvtv, ax, ay)
    tip = np.zeros(1, dtype=host.TIP_DTYPE)
    tip[0]["x"], tip[0]["y"] = 70.4, 58.6
    one = torch.tensor([1], dtype=torch.int32, device="cuda")
    tvec = torch.as_tensor(np.frombuffer(tip.tobytes(), dtype=np.uint8).copy()).cuda()
    I_tip = host.sr_integrals(p, du, dv, dvtu, dvtv, ax, ay, tip_count=one, tip_vector=tvec, count=3)
    uo, vo = torch.empty_like(du), torch.empty_like(dv)
    ax2, ay2 = zeros_like_rows(ny, nx), zeros_like_rows(ny, nx)
    host.advect_bfecc_cphi(p, du, dv, uo, vo, c, phi, adv_x=ax2, adv_y=ay2)
    assert torch.equal(ax, ax2) and torch.equal(ay, ay2)
    S = [zeros_like_rows(ny, nx) for _ in range(6)]
    S0 = [zeros_like_rows(ny, nx) for _ in range(6)]
    host.slice_fields(p, du, dv, S, S0, ax, ay, scheme=2)
    # slabs
    H = 3
    rows_sum = {k: torch.zeros(12 * host.sr_disc_slots(p), dtype=torch.float64, device="cuda") for k in ("c0", "tip")}
    for r in range(world):
        l = SlabLayout(ny, world, r, H)
        q = l.local_params(p)
        assert host.sr_disc_slots(q) == host.sr_disc_slots(p)
        sl = slice(l.g0, l.g1)
        lu, lv, lvtu, lvtv = dev(u[sl]), dev(v[sl]), dev(vtu[sl]), dev(vtv[sl])
        lax, lay_ = zeros_like_rows(l.ny_local, nx), zeros_like_rows(l.ny_local, nx)
        host.cxy_field(q, lax, lay_, c, phi)
        assert torch.equal(lax, ax[sl]) and torch.equal(lay_, ay[sl]), r
        ls = [torch.full((l.ny_local, nx), 7.0, dtype=torch.float64, device="cuda") for _ in range(6)]
        ls0 = [torch.full((l.ny_local, nx), 7.0, dtype=torch.float64, device="cuda") for _ in range(6)]
        host.slice_fields(q, lu, lv, ls, ls0, lax, lay_, scheme=2)       # slice_kernel accepts a slab as is
        for k6 in range(6):
            assert torch.equal(ls[k6][l.own_lo:l.own_hi], S[k6][l.j0:l.j1]), (r, k6)
            assert torch.equal(ls0[k6][l.own_lo:l.own_hi], S0[k6][l.j0:l.j1]), (r, k6)
        rows = torch.full((12 * host.sr_disc_slots(p),), 9.0, dtype=torch.float64, device="cuda")
        host.sr_integral_rows(q, lu, lv, lvtu, lvtv, lax, lay_, (p.tipx0, p.tipy0), (l.own_lo, l.own_hi), rows)
        rows_sum["c0"] += rows
        host.sr_integral_rows(q, lu, lv, lvtu, lvtv, lax, lay_, (float(tip[0]["x"]), float(tip[0]["y"])),
                              (l.own_lo, l.own_hi), rows)
        rows_sum["tip"] += rows
        luo, lvo = torch.full_like(lu, 5.0), torch.full_like(lv, 5.0)
        lax2, lay2 = zeros_like_rows(l.ny_local, nx), zeros_like_rows(l.ny_local, nx)
        host.advect_bfecc_cphi_rows(q, lu, lv, luo, lvo, c, phi, (l.own_lo, l.own_hi), adv_x=lax2, adv_y=lay2)
        own = slice(l.own_lo, l.own_hi)
        assert torch.equal(luo[own], uo[l.j0:l.j1]) and torch.equal(lvo[own], vo[l.j0:l.j1]), r
        assert torch.equal(lax2[own], ax[l.j0:l.j1]) and torch.equal(lay2[own], ay[l.j0:l.j1]), r
        if l.own_lo > 0:        # rows outside [row0, row1) are left alone
            assert (luo[:l.own_lo] == 5.0).all() and (lax2[:l.own_lo] == 0.0).all()
        if l.own_hi < l.ny_local:
            assert (luo[l.own_hi:] == 5.0).all() and (lax2[l.own_hi:] == 0.0).all()
    assert np.array_equal(host.sr_integrals_close(p, rows_sum["c0"]), I_whole)
    assert np.array_equal(host.sr_integrals_close(p, rows_sum["tip"]), I_tip)
    assert np.abs(I_whole).max() > 0 and not np.array_equal(I_whole, I_tip)


def _sr_case():
    nx = ny = 128
    u0 = wavy(nx, ny) - 0.05
    v0 = 0.3 * wavy(nx, ny, 0.21, 0.17, 1.3)
    return nx, ny, u0, v0


@pytest.mark.parametrize("world", [1, 2, 3])
def test_symmetry_reduction_steps_on_row_slabs(oracle, yh, world):
    """C3 (scaled down) on `world` row slabs == yh_sim_run_sr on the whole sheet, bit for bit.  The
    integration disc (radius 40 around a tip near (60, 66)) straddles every slab boundary."""
    nx, ny, u0, v0 = _sr_case()
    p = oracle.params_default(nx, ny, reduce_sym=True, tipOffsetX=40, tipOffsetY=40, tipx0=60.0, tipy0=66.0)
    nsteps = 14
    sim = yh.Sim(p)
    sim.set_state(u0[None], v0[None])
    rec = sim.run_sr(nsteps)
    gu, gv = sim.get_state()
    gc, gphi = sim.sr_state()
    sim.close()
    device = torch.device("cuda", torch.cuda.current_device())
    runners = [SlabRunner(p, rank=r, world=world, halo=p.timeIntOrder + 3, device=device, transport="nccl")
               for r in range(world)]
    records = [[] for _ in runners]
    for r in runners:
        r.load_global(u0, v0)
        r.sr_setup()
    advance_sr_emulated(runners, nsteps, records)
    torch.cuda.synchronize()
    got_u = np.concatenate([r.owned()[0].cpu().numpy() for r in runners])
    got_v = np.concatenate([r.owned()[1].cpu().numpy() for r in runners])
    assert np.array_equal(got_u, gu[0]) and np.array_equal(got_v, gv[0])
    for r, rc in zip(runners, records):
        assert np.array_equal(np.array(rc), rec), r.rank           # every rank holds the same (c, phi) history
        assert np.array_equal(np.array(r.c), gc) and np.array_equal(np.array(r.phi), gphi)
    assert np.abs(rec[:, :3]).max() > 1e-3, "the drift must be non-trivial"
    if world > 1:   # and the tips of the last step concatenate to the single-runner list
        one = SlabRunner(p, rank=0, world=1, halo=p.timeIntOrder + 3, device=device, transport="nccl")
        one.load_global(u0, v0)
        one.sr_setup()
        one.advance_sr(nsteps)
        cat = np.concatenate([r.sr_tips for r in runners])
        assert cat.tobytes() == one.sr_tips.tobytes()
