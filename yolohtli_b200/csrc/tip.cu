// tip.cu -- spiral-tip tracking (tipTracker.cu:20-611): intersection of the iso-lines
// u(t) = Uth and u(t+dt) = Uth inside each grid cell, appended to a device list.
//
// The reference appends with one atomicAdd per hit, so list order (and therefore "the last
// tip", which slice/trapz use as the disc centre) depends on scheduling.  Here the append is
// an ORDERED atomic compaction: CTAs take a ticket (atomicAdd), handle the ticket-th run of
// 1024 consecutive cells, and chain their hit counts by decoupled look-back, so the list
// comes out in ascending linear cell index, '+' root before '-' root, in one launch, with no
// sort and no host round trip.  Cost: 16 B per cell of reads -- HBM-bound.
#include "yh_common.cuh"
#include "yh_ordered.cuh"

namespace {

constexpr int TIP_THREADS = 256;
constexpr int TIP_CPT = 4;                       // consecutive cells per thread
constexpr int TIP_CHUNK = TIP_THREADS * TIP_CPT;

struct TipArgs {
  const double *past, *present;
  uint8_t *plot;
  int *count;
  yh_tip *vec;
  int capacity;
  int algorithm;
  float t;
  YhOrdered ord;               // ticket + look-back words (yh_ordered.cuh)
  const double *step;          // device-resident step counter (graph-replayable launches) or NULL
  double dt;
  int steps_ahead;
};

__device__ __forceinline__ bool equals_tol(double a, double b, double tol) {   // helper_functions.cu:58
  return (a == b) || ((a <= (b + tol)) && (a >= (b - tol)));
}

// gradient(), tipTracker.cu:519-566 (indices as shipped)
__device__ void tip_gradient(const YhK &k, int i, int j, double s, double t, const double *g,
                             float &gx, float &gy) {
  const int nx = k.nx, ny = k.nyg;   // j is a GLOBAL row; the arrays hold rows [jg0, jg0 + k.ny)
#define ID(a, b) ((a) + nx * ((b) - k.jg0))
  int S = (j > 0) ? ID(i, j - 1) : ID(i, j + 1);
  int Sx = ((j > 0) && (i < (nx - 1))) ? ID(i + 1, j - 1) : ID(i - 1, j + 1);
  int Sy = ID(i, j);
  int Sxy = (i < (nx - 1)) ? ID(i + 1, j) : ID(i - 1, j);
  int N = (j < (ny - 1)) ? ID(i, j + 1) : ID(i, j - 1);
  int Nx = ((i < (nx - 1)) && (j < (ny - 1))) ? ID(i + 1, j + 1) : ID(i - 1, j - 1);
  int Ny = (j < (ny - 2)) ? ID(i, j + 2) : ((j == (ny - 2)) ? ID(i, j) : ID(i, j - 1));
  int Nxy = ((i < (nx - 1)) && (j < (ny - 2))) ? ID(i + 1, j + 2)
          : ((j == (ny - 2)) ? ID(i - 1, j) : ID(i - 1, j - 1));
  int W = ID(i - 1, j);
  int Wx = ID(i, j);
  int Wy = ((i > 0) && (j < (ny - 1))) ? ID(i - 1, j + 1) : ID(i + 1, j - 1);
  int Wxy = (j < (ny - 1)) ? ID(i, j + 1) : ID(i - 1, j);
  int E = (i < (nx - 1)) ? ID(i + 1, j) : ID(i - 1, j);
  int Ex = (i < (nx - 2)) ? ID(i + 2, j) : ((i == (nx - 2)) ? ID(i, j) : ID(i - 1, j));
  int Ey = ((i < (nx - 1)) && (j < (ny - 1))) ? ID(i + 1, j + 1) : ID(i - 1, j - 1);
  int Exy = ((i < (nx - 2)) && (j < (ny - 1))) ? ID(i + 2, j + 1)
          : ((i == (nx - 2)) ? ID(i, j - 1) : ID(i - 1, j - 1));
#undef ID
  double gx1 = (g[E] - g[W]) * k.invdx, gy1 = (g[N] - g[S]) * k.invdy;
  double gx2 = (g[Ex] - g[Wx]) * k.invdx, gy2 = (g[Nx] - g[Sx]) * k.invdy;
  double gx3 = (g[Ey] - g[Wy]) * k.invdx, gy3 = (g[Ny] - g[Sy]) * k.invdy;
  double gx4 = (g[Exy] - g[Wxy]) * k.invdx, gy4 = (g[Nxy] - g[Sxy]) * k.invdy;
  gx = (float)((1.0 - s) * (1.0 - t) * gx1 + s * (1.0 - t) * gx2 + t * (1.0 - s) * gx3 + s * t * gx4);
  gy = (float)((1.0 - s) * (1.0 - t) * gy1 + s * (1.0 - t) * gy2 + t * (1.0 - s) * gy3 + s * t * gy4);
}

// Up to two candidate roots of one cell; returns how many pass tipRecordPlane's test.
// (i, j): j is the GLOBAL row (row slabs hold rows [jg0, jg0 + k.ny); whole sheet: jg0 = 0).
__device__ int tip_cell(const YhK &k, const TipArgs &a, int i, int j, float2 *roots) {
  const int nx = k.nx, ny = k.nyg;
  bool inside;
  if (k.solidSwitch) {   // tipTracker.cu:49-51
    int ic = i - nx / 2, jc = j - ny / 2;
    inside = (ic * ic + jc * jc) < k.tipOffX * k.tipOffY;
  } else if (a.algorithm == 2) {   // :347-348
    inside = (i >= (nx / 2 - k.tipOffX)) && (i < (nx / 2 + k.tipOffX)) &&
             (j >= (ny / 2 - k.tipOffY)) && (j < (ny / 2 + k.tipOffY));
  } else {   // :126
    inside = (i >= 1) && (i < (nx - 2)) && (j >= 1) && (j < (ny - 2));
  }
  if (!inside) return 0;
  const int s0 = i + nx * (j - k.jg0);
  const int sx = (i < (nx - 1)) ? s0 + 1 : s0;
  const int sy = (j < (ny - 1)) ? s0 + nx : s0;
  const int sxy = ((j < (ny - 1)) && (i < (nx - 1))) ? s0 + nx + 1 : s0;
  const double x1 = a.present[s0], x2 = a.present[sx], x4 = a.present[sy], x3 = a.present[sxy];
  const double y1 = a.past[s0], y2 = a.past[sx], y4 = a.past[sy], y3 = a.past[sxy];
  const double Uth = k.Uth;
  int n = 0;
  if (a.algorithm == 1) {   // :150-200
    const double x3y1 = x3 * y1, x4y1 = x4 * y1, x3y2 = x3 * y2, x4y2 = x4 * y2;
    const double x1y3 = x1 * y3, x2y3 = x2 * y3, x1y4 = x1 * y4, x2y4 = x2 * y4;
    const double x2y1 = x2 * y1, x1y2 = x1 * y2, x4y3 = x4 * y3, x3y4 = x3 * y4;
    const double den1 = 2.0 * (x3y1 - x4y1 - x3y2 + x4y2 - x1y3 + x2y3 + x1y4 - x2y4);
    const double den2 = 2.0 * (x2y1 - x3y1 - x1y2 + x4y2 + x1y3 - x4y3 - x2y4 + x3y4);
    const double ctn1 = x1 - x2 + x3 - x4 - y1 + y2 - y3 + y4;
    const double ctn2 = x3y1 - 2.0 * x4y1 + x4y2 - x1y3 + 2.0 * x1y4 - x2y4;
    const double disc = sqrt(4.0 * (x3y1 - x3y2 - x4y1 + x4y2 - x1y3 + x1y4 + x2y3 - x2y4)
                                 * (x4y1 - x1y4 + Uth * (x1 - x4 - y1 + y4))
                             + (-ctn2 + Uth * ctn1) * (-ctn2 + Uth * ctn1));
    const double px = ctn2 - Uth * ctn1;
    const double py = Uth * ctn1 - x3y1 + x4y2 + x1y3 - x2y4 + 2.0 * (x2y1 - x1y2);
    const bool ok = k.solidSwitch ? true : (disc >= 0.0);
    // The root is (N/D) cast to float and must land in (0, 1) (:182-185, :215).  N == 0, opposite
    // signs, or |N| >= |D| put the exact quotient outside (0, 1) and rounding is monotone, so those
    // candidates are rejected WITHOUT the FP64 division (4 per cell, most of this kernel's
    // arithmetic); NaNs fail every comparison here and fall through to the literal path.
    auto outside = [](double N, double D) {
      return (N == 0.0) || (N > 0.0 && D < 0.0) || (N < 0.0 && D > 0.0) || (fabs(N) >= fabs(D));
    };
#pragma unroll
    for (int sg = 0; sg < 2; sg++) {   // '+' root first (:182-183), then '-'
      const double Nx = sg ? px - disc : px + disc, Ny = sg ? py - disc : py + disc;
      if (!ok || outside(Nx, den1) || outside(Ny, den2)) continue;
      float2 r;
      r.x = (float)(Nx / den1); r.y = (float)(Ny / den2);
      if (((r.x > 0.0) && (r.x < 1.0)) && ((r.y > 0.0) && (r.y < 1.0))) roots[n++] = r;
    }
  } else {   // Newton, :384-424
    double s = 0.5, t = 0.5;
    for (int it = 0; it < 4; it++) {
      const double r1 = x1 * (1.0 - s) * (1.0 - t) + x2 * s * (1.0 - t) + x3 * s * t + x4 * (1.0 - s) * t - Uth;
      const double r2 = y1 * (1.0 - s) * (1.0 - t) + y2 * s * (1.0 - t) + y3 * s * t + y4 * (1.0 - s) * t - Uth;
      const double J11 = -x1 * (1.0 - t) + x2 * (1.0 - t) + x3 * t - x4 * t;
      const double J21 = -y1 * (1.0 - t) + y2 * (1.0 - t) + y3 * t - y4 * t;
      const double J12 = -x1 * (1.0 - s) - x2 * s + x3 * s + x4 * (1.0 - s);
      const double J22 = -y1 * (1.0 - s) - y2 * s + y3 * s + y4 * (1.0 - s);
      const double detJ = J11 * J22 - J12 * J21;
      if (!equals_tol(detJ, 0.0, 1e-14)) {
        const double sn = s - (J22 * r1 - J12 * r2) / detJ;
        const double tn = t - (-J21 * r1 + J11 * r2) / detJ;
        s = fmin(fmax(sn, 0.0), 1.0);
        t = fmin(fmax(tn, 0.0), 1.0);
      } else { s = -1.0; t = -1.0; }
    }
    const bool in01 = (s >= 0.0) && (s <= 1.0) && (t >= 0.0) && (t <= 1.0);
    const double u1 = in01 ? x1 * (1 - s) * (1.0 - t) + x2 * s * (1.0 - t) + x3 * s * t + x4 * (1.0 - s) * t : 0.0;
    const double u2 = in01 ? y1 * (1 - s) * (1.0 - t) + y2 * s * (1.0 - t) + y3 * s * t + y4 * (1.0 - s) * t : 0.0;
    if (equals_tol(u1, Uth, 1e-15) && equals_tol(u2, Uth, 1e-15)) {
      float2 r = make_float2((float)s, (float)t);
      if (((r.x > 0.0) && (r.x < 1.0)) && ((r.y > 0.0) && (r.y < 1.0))) roots[n++] = r;
    }
  }
  return n;
}

__global__ void __launch_bounds__(TIP_THREADS)
tip_kernel(const __grid_constant__ YhK k, const __grid_constant__ TipArgs a) {
  __shared__ int s_chunk;
  __shared__ int s_warp[TIP_THREADS / 32];
  __shared__ int s_base;
  const int tid = threadIdx.x;
  // graph-replayable launches take the time tag from the device-resident step counter
  const YhOrdered &ord = a.ord;
  float t_tag = a.t;
  if (a.step) t_tag = (float)(a.dt * (*a.step + (double)a.steps_ahead));
  if (tid == 0) s_chunk = (int)atomicAdd(&ord.state[0], 1ull);   // ticket = chunk, in launch order
  __syncthreads();
  const int chunk = s_chunk;
  const long long ncell = (long long)k.nx * (k.row1 - k.row0);   // local rows [row0, row1)

  float2 roots[TIP_CPT][2];
  int nr[TIP_CPT];
  int mine = 0;
#pragma unroll
  for (int r = 0; r < TIP_CPT; r++) {
    const long long cell = (long long)chunk * TIP_CHUNK + (long long)tid * TIP_CPT + r;
    nr[r] = 0;
    if (cell < ncell) {
      const int i = (int)(cell % k.nx), jl = k.row0 + (int)(cell / k.nx), j = jl + k.jg0;
      if (a.algorithm == 3) {   // abouzarTip_kernel, :434-517 -- raster only, no list
        const int s0 = i + k.nx * jl;
        const int sx = (i < (k.nx - 1)) ? s0 + 1 : s0;
        const int sy = (j < (k.nyg - 1)) ? s0 + k.nx : s0;
        const int sxy = ((j < (k.nyg - 1)) && (i < (k.nx - 1))) ? s0 + k.nx + 1 : s0;
        const double v0 = a.present[s0], vx = a.present[sx], vy = a.present[sy], vxy = a.present[sxy];
        int s = (0.0 >= v0 - k.Uth) + (0.0 >= vx - k.Uth) + (0.0 >= vy - k.Uth) + (0.0 >= vxy - k.Uth);
        const bool bv = (s > 0) && (s < 4);
        s = (0.0 >= v0 - a.past[s0]) + (0.0 >= vx - a.past[sx]) + (0.0 >= vy - a.past[sy]) +
            (0.0 >= vxy - a.past[sxy]);
        const bool bdv = (s > 0) && (s < 4);
        if (a.plot && bdv && bv) a.plot[s0] = 1;
      } else {
        nr[r] = tip_cell(k, a, i, j, roots[r]);
      }
    }
    mine += nr[r];
  }

  // block-wide exclusive scan of `mine` in thread order (= cell order), then the chunk prefix
  int total;
  const int excl = yh_block_excl_scan<TIP_THREADS>(mine, s_warp, total);
  if (tid < 32) {   // warp 0 chains this chunk to its predecessors
    const unsigned prefix = yh_ordered_prefix(ord, chunk, (unsigned)total, a.count);
    if (tid == 0) s_base = (int)prefix;
  }
  __syncthreads();
  if (mine == 0) return;

  int pos = s_base + excl;
#pragma unroll
  for (int r = 0; r < TIP_CPT; r++) {
    if (nr[r] == 0) continue;
    const long long cell = (long long)chunk * TIP_CHUNK + (long long)tid * TIP_CPT + r;
    const int i = (int)(cell % k.nx), j = k.row0 + (int)(cell / k.nx) + k.jg0;
    for (int q = 0; q < nr[r]; q++, pos++) {   // tipRecordPlane, :210-241
      const float2 tip = roots[r][q];
      float gx = 0.f, gy = 0.f;
      if (k.tipGrad) tip_gradient(k, i, j, tip.x, tip.y, a.present, gx, gy);
      yh_tip d;
      d.x = (float)(i + tip.x); d.y = (float)(j + tip.y); d.vx = gx; d.vy = gy; d.t = t_tag;
      if (pos < a.capacity) a.vec[pos] = d;
      if (a.plot) {   // plot_field, helper_functions.cu:45-51
        const int xi = (int)floorf(d.x), yi = (int)floorf(d.y);
        a.plot[xi + k.nx * (yi - k.jg0)] = 1;
      }
    }
  }
}

}  // namespace

// Cells of local rows [row0, row1); a cell reads the row above it, so the last row must be the
// sheet's last row or have a (ghost) row after it.  Coordinates in the list are GLOBAL.
extern "C" int yh_tip_track_rows(const yh_params *p, const double *u_past, const double *u_present,
                                 uint8_t *tip_plot, int *tip_count, yh_tip *tip_vector, int capacity,
                                 double physical_time, int algorithm, int row0, int row1, void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  YH_REQUIRE(p && u_past && u_present && tip_count && tip_vector, "null pointer");
  YH_REQUIRE(algorithm >= 1 && algorithm <= 3, "tip algorithm must be 1, 2 or 3");
  YH_REQUIRE(capacity >= 0, "negative capacity");
  YH_REQUIRE(p->nx >= 4 && p->ny >= 1 && p->jg0 >= 0 && p->jg0 + p->ny <= p->ny_global, "bad slab");
  YH_REQUIRE(row0 >= 0 && row0 < row1 && row1 <= p->ny, "bad row range");
  YH_REQUIRE(row1 < p->ny || p->jg0 + p->ny == p->ny_global, "the last row needs the row after it");
  YH_REQUIRE(!p->tipGrad || (p->jg0 == 0 && p->ny == p->ny_global), "tipGrad reads up to two rows away: whole sheets only");
  YhK k = yh_make_k(p);
  k.row0 = row0; k.row1 = row1;
  const long long ncell = (long long)p->nx * (row1 - row0);
  const int nchunks = (int)((ncell + TIP_CHUNK - 1) / TIP_CHUNK);
  unsigned long long *state = nullptr;
  rc = yh_workspace(((size_t)nchunks + 1) * sizeof(unsigned long long), (void **)&state, 1);
  if (rc != YH_OK) return rc;
  const unsigned epoch = yh_next_epoch();
  TipArgs a{u_past, u_present, tip_plot, tip_count, tip_vector, capacity, algorithm,
            (float)physical_time, {state, epoch, nchunks}, nullptr, 0.0, 0};
  tip_kernel<<<nchunks, TIP_THREADS, 0, (cudaStream_t)stream>>>(k, a);
  YH_LAUNCH_CHECK();
  return YH_OK;
}

extern "C" int yh_tip_track(const yh_params *p, const double *u_past, const double *u_present,
                            uint8_t *tip_plot, int *tip_count, yh_tip *tip_vector, int capacity,
                            double physical_time, int algorithm, void *stream) {
  YH_REQUIRE(p != nullptr, "null pointer");
  YH_REQUIRE(p->jg0 == 0 && p->ny_global == p->ny, "whole sheets only (row slabs: yh_tip_track_rows)");
  return yh_tip_track_rows(p, u_past, u_present, tip_plot, tip_count, tip_vector, capacity, physical_time,
                           algorithm, 0, p->ny, stream);
}

// Same pass for the device-resident SR loop, replayable from a CUDA graph: the time tag is
// dt*(sr_state[YH_SR_STEP] + steps_ahead), read on the device, and the look-back words are cleared by
// a memset in front of the launch instead of being told apart by a per-launch epoch.
int yh_tip_track_device_step(const yh_params *p, const double *u_past, const double *u_present,
                             int *tip_count, yh_tip *tip_vector, int capacity, const double *sr_state,
                             int steps_ahead, cudaStream_t st) {
  YhK k = yh_make_k(p);
  const long long ncell = (long long)p->nx * p->ny;
  const int nchunks = (int)((ncell + TIP_CHUNK - 1) / TIP_CHUNK);
  unsigned long long *state = nullptr;
  int rc = yh_workspace(((size_t)nchunks + 1) * sizeof(unsigned long long), (void **)&state, 1);
  if (rc != YH_OK) return rc;
  YH_CUDA(cudaMemsetAsync(state + 1, 0, (size_t)nchunks * sizeof(unsigned long long), st));
  TipArgs a{u_past, u_present, nullptr, tip_count, tip_vector, capacity, p->tipAlgorithm, 0.f,
            {state, 0xABCDEFu, nchunks}, sr_state + YH_SR_STEP, p->dt, steps_ahead};
  tip_kernel<<<nchunks, TIP_THREADS, 0, st>>>(k, a);
  YH_LAUNCH_CHECK();
  return YH_OK;
}
