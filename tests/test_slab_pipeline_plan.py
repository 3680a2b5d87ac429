"""CPU: the schedule of the pipelined yh_slab_run_host (csrc/slab.cu) as host arithmetic.

The copies of a slab's rows are hidden behind the time steps by running block (c, b) -- chunk c from level b-1 to
level b, a level = one block of n time steps needing h rows of the previous level on either side -- as soon as
chunk c has ARRIVED.  That is only correct if the regions obey a few invariants, checked here for many sheets,
partitions and run lengths through yh_slab_pipeline_plan / yh_slab_pipeline_region (no device involved):

  1. every rank of a partition derives the same plan (block length, halo per block, levels, chunks);
  2. at every level the regions of a slab are non-empty, ordered, contiguous and cover the owned rows minus the
     receding edges that face a neighbour (the wedges caught up later with one halo exchange per level);
  3. block (c, b) reads only rows that hold level b-1 and belong to chunks that arrived no later than c: inside the
     level b-1 regions of chunks c-1 and c (level 0: inside the rows uploaded so far), never a ghost row;
  4. it overwrites (same ping-pong half) level b-2 rows that no later block still reads: block (c+1, b-1), which
     runs AFTER (c, b) in the chunk-major order, starts reading exactly where (c, b) stops writing;
  5. the wedge rows [lo, lo + h*j) are caught up from level j-1 rows that nobody has overwritten, and at the last
     level regions + wedges tile the owned rows exactly (what goes back to the host)."""
import ctypes as C

import pytest

from yolohtli_b200 import _lib


def plan(l, ny, world, rank, halo, K, fast, nsteps, tb=0):
    out = (C.c_int * 8)()
    P = l.yh_slab_pipeline_plan(ny, world, rank, halo, K, int(fast), nsteps, tb, out)
    return P, dict(zip(("n", "h", "B", "P", "C", "S", "lo", "hi"), out[:]))


def region(l, ny, world, rank, halo, K, fast, nsteps, c, b, tb=0):
    rows = (C.c_int * 2)()
    assert l.yh_slab_pipeline_region(ny, world, rank, halo, K, int(fast), nsteps, tb, c, b, rows) == 0
    return rows[0], rows[1]


CASES = [(16384, 1, 4, 1, True, 1024), (16384, 8, 4, 1, True, 1024), (16384, 2, 4, 1, True, 4096), (1024, 2, 4, 1, True, 203),
         (1200, 3, 4, 1, True, 171), (1024, 4, 4, 1, True, 150), (1100, 3, 4, 4, False, 47), (8192, 8, 4, 4, False, 64),
         (4099, 5, 4, 2, False, 300), (16384, 8, 7, 1, True, 777), (2048, 2, 2, 1, True, 90), (40000, 7, 4, 1, True, 5000)]


@pytest.mark.parametrize("ny,world,halo,K,fast,nsteps", CASES)
def test_pipeline_plan_invariants(ny, world, halo, K, fast, nsteps):
    l = _lib.lib()
    plans = [plan(l, ny, world, r, halo, K, fast, nsteps) for r in range(world)]
    assert all(p[0] > 0 for p in plans), "these cases are meant to take the pipelined schedule"
    # 1. one plan for all ranks
    for key in ("n", "h", "B", "P", "C"):
        assert len({p[1][key] for p in plans}) == 1, key
    for rank, (P, pl) in enumerate(plans):
        n, h, Cn, S, lo, hi = pl["n"], pl["h"], pl["C"], pl["S"], pl["lo"], pl["hi"]
        up, down = rank > 0, rank < world - 1
        assert h == n * K and h <= halo and 2 * P + 1 <= pl["B"]
        X = [lo + c * S for c in range(Cn)] + [hi]              # chunk c arrives as rows [X[c], X[c+1])
        assert all(X[c] < X[c + 1] for c in range(Cn))
        reg = {(c, b): region(l, ny, world, rank, halo, K, fast, nsteps, c, b) for c in range(Cn) for b in range(1, P + 1)}
        for b in range(1, P + 1):
            # 2. ordered, contiguous, non-empty, covering the owned rows minus the receding edges
            assert reg[(0, b)][0] == (lo + h * b if up else lo)
            assert reg[(Cn - 1, b)][1] == (hi - h * b if down else hi)
            for c in range(Cn):
                r0, r1 = reg[(c, b)]
                assert r1 - r0 >= 8, (c, b, r0, r1)
                if c:
                    assert r0 == reg[(c - 1, b)][1]
                # 3. reads [r0 - h, r1 + h) of level b-1, from chunks that arrived no later than c
                rd0 = r0 - h if (c > 0 or up) else r0            # the sheet's own edge mirrors, it reads nothing beyond
                rd1 = r1 + h if (c < Cn - 1 or down) else r1
                if b == 1:
                    assert rd0 >= lo and rd1 <= X[c + 1], "level 0 = rows uploaded so far, no ghost row"
                else:
                    first = reg[(max(c - 1, 0), b - 1)][0]
                    assert rd0 >= first and rd1 <= reg[(c, b - 1)][1]
                    assert rd0 >= lo and rd1 <= hi
                # 4. block (c+1, b-1) runs later and still reads level b-2 in the half that (c, b) overwrites
                if b >= 2 and c + 1 < Cn:
                    nxt0 = reg[(c + 1, b - 1)][0]
                    assert nxt0 - h >= r1, "a later block would read rows this block has overwritten"
        # 5. wedges: level j of the top wedge reads rows up to lo + h*(j+1) of level j-1; the chunk-0 blocks of levels
        #    j+1, j+3, ... (same half) start at lo + h*(j+1) or below it, never inside
        if up:
            for j in range(1, P + 1):
                need_hi = lo + h * (j + 1)
                for b in range(j + 1, P + 1, 2):
                    assert reg[(0, b)][0] >= need_hi
        if down:
            for j in range(1, P + 1):
                need_lo = hi - h * (j + 1)
                for b in range(j + 1, P + 1, 2):
                    assert reg[(Cn - 1, b)][1] <= need_lo
        # what goes home: regions of the last level + the two wedges = the owned rows, each row once
        cover = [reg[(c, P)] for c in range(Cn)]
        if up:
            cover.insert(0, (lo, lo + h * P))
        if down:
            cover.append((hi - h * P, hi))
        assert cover[0][0] == lo and cover[-1][1] == hi
        assert all(cover[q][1] == cover[q + 1][0] for q in range(len(cover) - 1))


def test_pipeline_plan_falls_back():
    l = _lib.lib()
    assert plan(l, 512, 1, 0, 4, 1, True, 8)[0] == 0          # too few blocks
    assert plan(l, 200, 1, 0, 4, 1, True, 1000)[0] == 0       # slab too small for four chunks
    assert plan(l, 16384, 1, 0, 2, 4, False, 1000)[0] == 0    # halo smaller than one RK4 step needs
    assert plan(l, 16384, 8, 3, 4, 1, True, 1024)[0] == 28    # the bench's end-to-end call at N = 8
    assert plan(l, 16384, 1, 0, 4, 1, True, 1024)[1]["C"] == 8
