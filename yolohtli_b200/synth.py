"""Synthetic inputs with fixed seeds: the reference ships neither its
obstacle masks (common/holes*.dat are absent, .MISSING_LARGE_BLOBS) nor saved initial
conditions, so both are regenerated from the recipes in the reference tree."""
import numpy as np


def cross_field_ic(nx, ny):
    """initGates, main.cu:606-618: u = 1 for i < nx/8, v = 1 for j >= ny/2."""
    u = np.zeros((ny, nx), dtype=np.float64)
    v = np.zeros((ny, nx), dtype=np.float64)
    u[:, : nx // 8] = 1.0
    v[ny // 2:, :] = 1.0
    return u, v


class XorShift64Star:
    def __init__(self, seed=0x9E3779B97F4A7C15):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next_u64(self):
        x = self.s
        x ^= x >> 12
        x ^= (x << 25) & 0xFFFFFFFFFFFFFFFF
        x ^= x >> 27
        self.s = x
        return (x * 0x2545F4914F6CDD1D) & 0xFFFFFFFFFFFFFFFF

    def uniform(self):
        return (self.next_u64() >> 11) * (1.0 / 9007199254740992.0)

    def randint(self, lo, hi):
        """uniform integer in [lo, hi]"""
        return lo + int(self.uniform() * (hi - lo + 1))


def fibrillation_ic(nx, ny, patch=128, seed=0x9E3779B97F4A7C15, rows=None):
    """Domain tiled in patch x patch squares, each a cross-field pair (a band of u = 1 and a
    half-patch of v = 1) rotated by a seeded choice of 0/90/180/270 degrees: many interacting
    spirals -> fibrillation-like activity.  rows=(a, b) returns only global rows [a, b) (each
    rank of a slab run builds just its own rows; the orientation table is drawn for the whole
    domain first, so every rank sees the same field)."""
    rng = XorShift64Star(seed)
    npy, npx = (ny + patch - 1) // patch, (nx + patch - 1) // patch
    rot = [[rng.next_u64() & 3 for _ in range(npx)] for _ in range(npy)]
    a0, b0 = rows if rows is not None else (0, ny)
    u = np.zeros((b0 - a0, nx), dtype=np.float64)
    v = np.zeros((b0 - a0, nx), dtype=np.float64)
    pu = np.zeros((patch, patch))
    pv = np.zeros((patch, patch))
    pu[:, : patch // 8] = 1.0
    pv[patch // 2:, :] = 1.0
    rots = [(np.rot90(pu, k), np.rot90(pv, k)) for k in range(4)]
    for pj in range(a0 // patch, (b0 - 1) // patch + 1):
        j0 = pj * patch
        lo, hi = max(j0, a0), min(j0 + patch, b0, ny)
        for pi in range(npx):
            i0 = pi * patch
            w = min(patch, nx - i0)
            a, b = rots[rot[pj][pi]]
            u[lo - a0:hi - a0, i0:i0 + w] = a[lo - j0:hi - j0, :w]
            v[lo - a0:hi - a0, i0:i0 + w] = b[lo - j0:hi - j0, :w]
    return u, v


def hole_mask(n, seed=0x9E3779B97F4A7C15, area=0.0184, rmin=0.1, rmax=0.8, a=-2.75):
    """common/Hole_generator.m:3-27: non-overlapping discs, radii r^2/0.2^2 px with r from a
    power law on [rmin, rmax], until the hole area fraction exceeds `area`; 1 = tissue."""
    rng = XorShift64Star(seed)
    holes = np.zeros((n, n), dtype=bool)
    yy, xx = np.mgrid[1:n + 1, 1:n + 1]
    e = 1.0 + a
    while holes.mean() <= area:
        cx, cy = rng.randint(1, n), rng.randint(1, n)   # draw order of Hole_generator.m:18-20
        xi = rng.uniform()
        r = (rmin ** e + xi * (rmax ** e - rmin ** e)) ** (1.0 / e)
        R = r * r / 0.04
        x0, x1 = max(1, int(cx - R) - 1), min(n, int(cx + R) + 1)
        y0, y1 = max(1, int(cy - R) - 1), min(n, int(cy + R) + 1)
        sub = np.sqrt((xx[y0 - 1:y1, x0 - 1:x1] - cx) ** 2 + (yy[y0 - 1:y1, x0 - 1:x1] - cy) ** 2) <= R
        if not sub.any() or (holes[y0 - 1:y1, x0 - 1:x1] & sub).any():
            continue
        holes[y0 - 1:y1, x0 - 1:x1] |= sub
    return (~holes).astype(np.uint8)


def write_mask_dat(path, mask):
    """The reference's holes<N>.dat: one %e float per line, file order = i + j*nx
    (main.cu:676-680 reads with fscanf "%f" and thresholds at 0.5)."""
    with open(path, "w") as f:
        for val in np.asarray(mask, dtype=np.float64).ravel():
            f.write("%e\n" % val)


def read_mask_dat(path, n):
    vals = np.loadtxt(path, dtype=np.float32)
    assert vals.size == n * n
    return (vals > 0.5).astype(np.uint8).reshape(n, n)


def stim_area_square(nx, ny):
    """domainObjects, main.cu:839: stimArea = (j >= 35) on the square domain."""
    m = np.ones((ny, nx), dtype=np.uint8)
    m[:35, :] = 0
    return m
