// sr.cu -- symmetry-reduction (co-moving frame) pieces of the step (main.cu:894-954):
//   yh_slice          slice_kernel                symmetryReduction.cu:72-222
//   yh_trapz          trapz_kernel x 12           integralTrapz.cu:17-184   (one pass, one D2H)
//   yh_sr_integrals   slice + trapz fused: tangent fields formed in registers, the 12
//                     nx*ny slice arrays (96 B/cell of writes that only trapz re-reads) never exist
//   yh_cxy_field      Cxy_field_kernel            symmetryReduction.cu:20-62
//   yh_advect_bfecc*  advFDBFECC_kernel           advFDBFECC.cu:18-353, three SYNCHRONOUS
//                     sweeps in a shared-memory tile with a 3-cell halo (no uf/ub/ue arrays)
//
// Reductions are warp-shuffle / fixed-order so the 12 integrals are deterministic: one warp
// per grid row (lane l sums i = l, l+32, ...; xor-butterfly 16..1), rows added in ascending j.
#include <math.h>

#include "yh_common.cuh"

namespace {

// Disc centre (symmetryReduction.cu:98-104 / integralTrapz.cu:42-48).  An empty list keeps
// (tipx0, tipy0) instead of reading tip_vector[-1] (defect B3).
__device__ __forceinline__ void disc_centre(const YhK &k, float tipx0, float tipy0, const int *tip_count,
                                            const yh_tip *tv, int count, int &cx, int &cy) {
  float fx = tipx0, fy = tipy0;
  if (count != 0 && tip_count) {
    const int n = *tip_count;
    if (n > 0) { fx = tv[n - 1].x; fy = tv[n - 1].y; }
  }
  cx = __float2int_rn(fx - (float)(k.nx / 2));
  cy = __float2int_rn(fy - (float)(k.nyg / 2));
}

// (i, j) with j a GLOBAL row: a process may hold only rows [jg0, jg0 + ny) of the sheet (row slabs,
// ghost rows included); mirror rules apply at the global edges.  Whole sheet: jg0 = 0, nyg = ny.
__device__ __forceinline__ int IDX(const YhK &k, int i, int j) { return i + k.nx * (j - k.jg0); }

__device__ __forceinline__ double fb2x(const YhK &k, const double *f, int i, int j, int C, int E, int W, double ax) {
  const int WW = IDX(k, yh_mir(i - 2, k.nx), j), EE = IDX(k, yh_mir(i + 2, k.nx), j);
  return (ax > 0.0) ? (-3.0 * f[C] + 4.0 * f[E] - f[EE]) * k.invdx : (3.0 * f[C] - 4.0 * f[W] + f[WW]) * k.invdx;
}
__device__ __forceinline__ double fb2y(const YhK &k, const double *f, int i, int j, int C, int N, int S, double ay) {
  const int SS = IDX(k, i, yh_mir(j - 2, k.nyg)), NN = IDX(k, i, yh_mir(j + 2, k.nyg));
  return (ay > 0.0) ? (-3.0 * f[C] + 4.0 * f[N] - f[NN]) * k.invdy : (3.0 * f[C] - 4.0 * f[S] + f[SS]) * k.invdy;
}
__device__ __forceinline__ double cen2x(const YhK &k, const double *f, int i, int j, int E, int W) {
  const int WW = IDX(k, yh_mir(i - 2, k.nx), j), EE = IDX(k, yh_mir(i + 2, k.nx), j);
  return (f[EE] - 8.0 * f[E] + 8.0 * f[W] - f[WW]) * k.invdx * (1.0 / 6.0);
}
__device__ __forceinline__ double cen2y(const YhK &k, const double *f, int i, int j, int N, int S) {
  const int SS = IDX(k, i, yh_mir(j - 2, k.nyg)), NN = IDX(k, i, yh_mir(j + 2, k.nyg));
  return (f[NN] - 8.0 * f[N] + 8.0 * f[S] - f[SS]) * k.invdy * (1.0 / 6.0);
}

// tangent values of one cell: s = upwind (slice), s0 = template (slice0); order ux uy ut vx vy vt
__device__ __forceinline__ void slice_cell(const YhK &k, const double *gu, const double *gv, const double *ax,
                                           const double *ay, int scheme, bool sc, int i, int j, bool want0,
                                           double s[6], double s0[6]) {
  const int c = IDX(k, i, j);
  const double x = (double)i;                                    // (i + nx*j) % nx
  const double y = (double)floorf((float)(j % k.nx));            // (idx / nx) % nx as shipped (:109-110)
  const int S = IDX(k, i, yh_mir(j - 1, k.nyg)), N = IDX(k, i, yh_mir(j + 1, k.nyg));
  const int W = IDX(k, yh_mir(i - 1, k.nx), j), E = IDX(k, yh_mir(i + 1, k.nx), j);
  const bool on = (scheme == 1) ? true : sc;
  if (on) {
    const double axc = ax[c], ayc = ay[c];
    s[0] = fb2x(k, gu, i, j, c, E, W, axc);
    s[1] = fb2y(k, gu, i, j, c, N, S, ayc);
    s[3] = fb2x(k, gv, i, j, c, E, W, axc);
    s[4] = fb2y(k, gv, i, j, c, N, S, ayc);
    s[2] = k.hx * x * s[1] - k.hy * y * s[0];
    s[5] = k.hx * x * s[4] - k.hy * y * s[3];
  } else {
#pragma unroll
    for (int q = 0; q < 6; q++) s[q] = 0.0;
  }
  if (!want0) return;
  if (scheme == 1) {   // :151-157
    s0[0] = s[0]; s0[1] = s[1]; s0[3] = s[3]; s0[4] = s[4];
    s0[2] = k.hx * x * s[1] - k.hy * y * s[0];
    s0[5] = k.hx * x * s[4] - k.hy * y * s[3];
  } else if (sc) {     // :191-202
    s0[0] = cen2x(k, gu, i, j, E, W);
    s0[1] = cen2y(k, gu, i, j, N, S);
    s0[3] = cen2x(k, gv, i, j, E, W);
    s0[4] = cen2y(k, gv, i, j, N, S);
    s0[2] = k.hx * x * s0[1] - k.hy * y * s0[0];
    s0[5] = k.hx * x * s0[4] - k.hy * y * s0[3];
  } else {
#pragma unroll
    for (int q = 0; q < 6; q++) s0[q] = 0.0;
  }
}

struct P6 { double *p[6]; };
struct CP6 { const double *p[6]; };

struct SliceArgs {
  const double *u, *v, *ax, *ay;
  P6 s, s0;
  int start, scheme, count;
  const int *tip_count;
  const yh_tip *tv;
  float tipx0, tipy0;
};

__global__ void __launch_bounds__(256)
slice_kernel(const __grid_constant__ YhK k, const __grid_constant__ SliceArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int jl = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= k.nx || jl >= k.ny) return;
  const int j = jl + k.jg0;
  int cx, cy;
  disc_centre(k, a.tipx0, a.tipy0, a.tip_count, a.tv, a.count, cx, cy);
  const int ic = i - k.nx / 2, jc = j - k.nyg / 2;
  const bool sc = ((ic - cx) * (ic - cx) + (jc - cy) * (jc - cy)) < k.tipOffX * k.tipOffY;
  double s[6], s0[6];
  slice_cell(k, a.u, a.v, a.ax, a.ay, a.scheme, sc, i, j, a.start != 0, s, s0);
  const int c = IDX(k, i, j);
#pragma unroll
  for (int q = 0; q < 6; q++) a.s.p[q][c] = s[q];
  if (a.start) {
#pragma unroll
    for (int q = 0; q < 6; q++) a.s0.p[q][c] = s0[q];
  }
}

// ---- the 12 integrals -------------------------------------------------------------------
struct IntArgs {
  // FUSED: from u, v, adv; otherwise from the slice arrays
  const double *u, *v, *ax, *ay;
  CP6 s, s0;
  const double *vtu, *vtv;
  double *rows;        // [nrows_max][12] row sums
  double *out;         // [12] device result
  unsigned *done;      // CTAs finished (the last one closes the pass and re-arms it)
  double *sr;          // device-resident (c, phi, cos, sin): solve on the device when non-NULL
  double *log_base;    // (c, phi) records of the run (device solve), may be NULL
  double step0;        // first step of the run: no deferred phi update before it
  double dt;
  int count;
  const int *tip_count;
  const yh_tip *tv;
  float tipx0, tipy0;
  int R;               // ceil(sqrt(tipOffX*tipOffY)): rows cy-R .. cy+R can intersect the disc
  int own0, own1;      // GLOBAL rows this process sums (row slabs; whole sheet: 0, ny); other slots get +0.0
};

// Summation order (shared with the oracle, yh_oracle.c): per grid row 256 accumulators -- warp w,
// lane l takes i = 32w + l, +256, ... -- lanes combined by the xor-butterfly 16..1, the 8 warp
// sums added in ascending w; then over the row slots: 32 partial sums (partial l takes slots
// l, l+32, ...), combined by the same butterfly.  One CTA per row slot: a 512-cell row is two cells
// per thread instead of sixteen per lane of a lone warp (the pass is pure latency).
// The LAST CTA to finish closes the pass in the same launch: the 12 totals and, for the
// device-resident SR step (yh_sim_run_sr_device), the deferred frame update phi += c*dt of the
// PREVIOUS step (main.cu:936-938), the (c, phi) record (main.cu:902-903), the 3x3 solve
// (symmetryReduction.cu:386-416) and cos/sin(phi.t) for the advection that follows.  cos/sin are
// libdevice's there and libm's in the host path: results agree to rounding, not bit for bit.
constexpr int ROW_WARPS = 8;

template <bool FUSED>
__global__ void __launch_bounds__(ROW_WARPS * 32)
integrals_rows_kernel(const __grid_constant__ YhK k, const __grid_constant__ IntArgs a) {
  __shared__ double s_w[ROW_WARPS][12];
  __shared__ double s_I[12];
  __shared__ bool s_last;
  const int slot = blockIdx.x;   // row slot
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int cx, cy;
  disc_centre(k, a.tipx0, a.tipy0, a.tip_count, a.tv, a.count, cx, cy);
  const int j = cy + k.nyg / 2 - a.R + slot;   // (global) grid row of this slot
  double acc[12];
#pragma unroll
  for (int q = 0; q < 12; q++) acc[q] = 0.0;
  if (j >= a.own0 && j < a.own1) {
    const int jc = j - k.nyg / 2;
    for (int i = threadIdx.x; i < k.nx; i += ROW_WARPS * 32) {
      const int ic = i - k.nx / 2;
      const bool sc = ((ic - cx) * (ic - cx) + (jc - cy) * (jc - cy)) < k.tipOffX * k.tipOffY;
      if (!sc) continue;   // the reference adds exactly +0.0 here (integralTrapz.cu:55-57)
      const int c = IDX(k, i, j);
      double s[6], s0[6];
      if (FUSED) {
        slice_cell(k, a.u, a.v, a.ax, a.ay, 2, true, i, j, true, s, s0);
      } else {
#pragma unroll
        for (int q = 0; q < 6; q++) { s[q] = a.s.p[q][c]; s0[q] = a.s0.p[q][c]; }
      }
      const double vu = a.vtu[c], vv = a.vtv[c];
#pragma unroll
      for (int m = 0; m < 3; m++) {
#pragma unroll
        for (int b = 0; b < 3; b++) acc[3 * m + b] += 4.0 * (s0[m] * s[b] + s0[m + 3] * s[b + 3]);
        acc[9 + m] += 4.0 * (s0[m] * vu + s0[m + 3] * vv);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 12; q++) {
    double x = acc[q];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) x = x + __shfl_xor_sync(0xffffffffu, x, m);
    if (lane == 0) s_w[wid][q] = x;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    double r = s_w[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < ROW_WARPS; w++) r = r + s_w[w][threadIdx.x];
    __stcg(a.rows + (size_t)slot * 12 + threadIdx.x, r);
    __threadfence();
  }
  if (!a.done) return;   // slab form: the row sums are combined across processes before the closing
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(a.done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;

  // ---- the last CTA: totals over the row slots (every row sum is visible: fence + atomic) ----
  __threadfence();
  const int nrows = gridDim.x;
  for (int q = wid; q < 12; q += ROW_WARPS) {   // one warp per integral
    double x = 0.0;
    for (int w = lane; w < nrows; w += 32) x += __ldcg(a.rows + (size_t)w * 12 + q);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) x = x + __shfl_xor_sync(0xffffffffu, x, m);
    if (lane == 0) {
      s_I[q] = 0.25 * k.hx * k.hy * x;   // integralTrapz.cu:79
      a.out[q] = s_I[q];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    *a.done = 0u;   // re-arm
    if (a.sr) {
      double c[3], phi[3];
      const double step = a.sr[YH_SR_STEP];
      const double dt_phi = (step > a.step0) ? a.dt : 0.0;
      for (int m = 0; m < 3; m++) { c[m] = a.sr[YH_SR_C + m]; phi[m] = a.sr[YH_SR_PHI + m] + c[m] * dt_phi; }
      if (a.log_base) {
        double *log_row = a.log_base + 6 * (size_t)(step - a.step0);
        for (int m = 0; m < 3; m++) { log_row[m] = c[m]; log_row[3 + m] = phi[m]; }
      }
      a.sr[YH_SR_STEP] = step + 1.0;
      const double cs = cos(phi[2]), sn = sin(phi[2]);
      yh_solve3(s_I, cs, sn, c);
      for (int m = 0; m < 3; m++) { a.sr[YH_SR_C + m] = c[m]; a.sr[YH_SR_PHI + m] = phi[m]; }
      a.sr[YH_SR_CS] = cs; a.sr[YH_SR_SN] = sn;
    }
  }
}

// Closing of the slab form: the totals over the row slots in the canonical order of the last-CTA
// code above (32 interleaved partial sums, xor-butterfly), one warp per integral.
__global__ void __launch_bounds__(12 * 32)
integrals_close_kernel(const double *rows, int nrows, double hx, double hy, double *out) {
  const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
  double x = 0.0;
  for (int w = lane; w < nrows; w += 32) x += rows[(size_t)w * 12 + q];
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) x = x + __shfl_xor_sync(0xffffffffu, x, m);
  if (lane == 0) out[q] = 0.25 * hx * hy * x;   // integralTrapz.cu:79, same operation order as above
}

__global__ void sr_flush_phi_kernel(double *sr, double dt_phi) {
  if (threadIdx.x < 3) sr[YH_SR_PHI + threadIdx.x] = sr[YH_SR_PHI + threadIdx.x] + sr[YH_SR_C + threadIdx.x] * dt_phi;
}

int disc_radius_rows(const yh_params *p) {
  const long long r2 = (long long)p->tipOffsetX * p->tipOffsetY;
  int R = (int)ceil(sqrt((double)r2));
  if (R > p->ny_global) R = p->ny_global;
  return R;
}

int fetch12(const double *out_d, double *integrals_host, cudaStream_t st) {
  static thread_local double *pinned = nullptr;
  if (!pinned) YH_CUDA(cudaMallocHost(&pinned, 12 * sizeof(double)));
  YH_CUDA(cudaMemcpyAsync(pinned, out_d, 12 * sizeof(double), cudaMemcpyDeviceToHost, st));
  YH_CUDA(cudaStreamSynchronize(st));   // the ONE host sync of the SR step (reference: 12 + malloc/free)
  for (int q = 0; q < 12; q++) integrals_host[q] = pinned[q];
  return YH_OK;
}

int run_integrals(const yh_params *p, IntArgs &a, bool fused, double *integrals_host, cudaStream_t st) {
  YhK k = yh_make_k(p);
  const int R = disc_radius_rows(p);
  a.R = R;
  a.own0 = p->jg0; a.own1 = p->jg0 + p->ny;
  const int nrows = 2 * R + 1;
  double *ws = nullptr;
  int rc = yh_workspace(((size_t)nrows * 12 + 12 + 1) * sizeof(double), (void **)&ws, 2);   // zero-filled when (re)allocated
  if (rc != YH_OK) return rc;
  // fixed places for the counter and the result: the row count changes with the disc radius, and a
  // counter that moved onto old row sums would not start at zero
  a.done = reinterpret_cast<unsigned *>(ws); a.out = ws + 1; a.rows = ws + 13;
  if (fused) integrals_rows_kernel<true><<<nrows, ROW_WARPS * 32, 0, st>>>(k, a);
  else integrals_rows_kernel<false><<<nrows, ROW_WARPS * 32, 0, st>>>(k, a);
  YH_LAUNCH_CHECK();
  if (!integrals_host) return YH_OK;   // device-resident closing was done by the last CTA
  return fetch12(a.out, integrals_host, st);
}

// ---- Cxy ----------------------------------------------------------------------------------
struct CxyArgs { double *ax, *ay; const uint8_t *solid; double cx, cy, ct, cs, sn; };

// adv of one cell (symmetryReduction.cu:45-56); cos/sin(phi.t) are computed on the HOST so the
// field is reproducible by the plain-C oracle bit for bit.
// j is the GLOBAL row (row slabs: arrays hold rows [jg0, jg0 + ny)).
__device__ __forceinline__ void cxy_cell(const YhK &k, const CxyArgs &a, int i, int j, double &ax, double &ay) {
  const double x = (double)i;                               // (i + nx*j) % nx
  const double y = (double)floorf((float)(j % k.nx));       // (idx / nx) % nx as shipped
  ax = k.hy * y * a.ct - a.cx * a.cs + a.cy * a.sn;
  ay = -k.hx * x * a.ct - a.cx * a.sn - a.cy * a.cs;
  if (k.solidSwitch && !a.solid[IDX(k, i, j)]) { ax = 0.0; ay = 0.0; }
}

__global__ void __launch_bounds__(256)
cxy_kernel(const __grid_constant__ YhK k, const __grid_constant__ CxyArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= k.nx || j >= k.ny) return;
  double ax, ay;
  cxy_cell(k, a, i, j + k.jg0, ax, ay);
  a.ax[i + k.nx * j] = ax;
  a.ay[i + k.nx * j] = ay;
}

// ---- BFECC --------------------------------------------------------------------------------
constexpr int BT = 32;            // output tile edge
constexpr int BH = 3;             // halo
constexpr int BP = BT + 2 * BH;   // 38
constexpr int BTHREADS = 512;
constexpr int FO = 3 * BP * BP;   // plane offset between the two fields: [g uf ue](u) [g uf ue](v) Rx Ry mask

struct BfArgs {
  const double *u_in, *v_in;
  double *u_out, *v_out;
  const double *ax, *ay;   // adv arrays (when !from_c)
  double *ax_out, *ay_out; // optional copy of the generated field (from_c)
  const uint8_t *solid;
  CxyArgs cxy;
  int from_c;
  const double *sr_state;   // from_c with (c, cos, sin) read from device memory (yh_sim_run_sr_device)
};

__device__ __forceinline__ int sgn(double x) { int t = x < 0.0 ? -1 : 0; return x > 0.0 ? 1 : t; }

// tile-local index of GLOBAL cell (gi, gj), mirrored into the domain first (defect B5)
__device__ __forceinline__ int TL(const YhK &k, int gi, int gj, int ti0, int tj0) {
  gi = yh_mir(gi, k.nx); gj = yh_mir(gj, k.nyg);
  return (gi - ti0) + BP * (gj - tj0);
}

struct NbrIdx { int W, E, S, N, i2dW, W2, i2dE, E2, i2dS, S2, i2dN, N2; bool sc; };

// index selection of the Neumann branches (advFDBFECC.cu:44-66 solid, :112-126 square)
__device__ __forceinline__ NbrIdx neumann_idx(const YhK &k, const uint8_t *msk, int gi, int gj, int ti0, int tj0) {
  NbrIdx x;
  const int C = (gi - ti0) + BP * (gj - tj0);
  bool sw, se, sn, ss;
  if (k.solidSwitch) {
    x.sc = msk[C];
    sw = (gi > 0) && msk[C - 1];
    se = (gi < k.nx - 1) && msk[C + 1];
    sn = (gj > 0) && msk[C - BP];            // "sn" is the mask at j-1 (:47)
    ss = (gj < k.nyg - 1) && msk[C + BP];    // "ss" is the mask at j+1 (:48)
  } else {
    x.sc = true;
    sw = gi > 0; se = gi < (k.nx - 1);
    ss = gj > 0;            // square branch tests the index itself: S uses (j>0), N uses (j<ny-1)
    sn = gj < (k.nyg - 1);
  }
  if (k.solidSwitch) {
    const int Wm = TL(k, gi - 1, gj, ti0, tj0), Em = TL(k, gi + 1, gj, ti0, tj0);
    const int Sm = TL(k, gi, gj - 1, ti0, tj0), Nm = TL(k, gi, gj + 1, ti0, tj0);
    x.W = x.sc ? (sw ? Wm : Em) : C;
    x.E = x.sc ? (se ? Em : Wm) : C;
    x.S = x.sc ? (ss ? Sm : Nm) : C;
    x.N = x.sc ? (sn ? Nm : Sm) : C;
    x.i2dW = x.sc ? (sw ? C : Em) : C;  x.W2 = x.sc ? (sw ? Wm : C) : C;
    x.i2dE = x.sc ? (se ? C : Wm) : C;  x.E2 = x.sc ? (se ? Em : C) : C;
    x.i2dS = x.sc ? (ss ? C : Nm) : C;  x.S2 = x.sc ? (ss ? Sm : C) : C;
    x.i2dN = x.sc ? (sn ? C : Sm) : C;  x.N2 = x.sc ? (sn ? Nm : C) : C;
  } else {
    x.W = sw ? C - 1 : C + 1;   x.E = se ? C + 1 : C - 1;
    x.S = ss ? C - BP : C + BP; x.N = sn ? C + BP : C - BP;
    x.i2dW = sw ? C : C + 1;    x.W2 = sw ? C - 1 : C;
    x.i2dE = se ? C : C - 1;    x.E2 = se ? C + 1 : C;
    x.i2dS = ss ? C : C + BP;   x.S2 = ss ? C - BP : C;
    x.i2dN = sn ? C : C - BP;   x.N2 = sn ? C + BP : C;
  }
  return x;
}

// Dirichlet neighbour values (advFDBFECC.cu:171-183 solid, :265-269 square).
// bnd = value used where the neighbour is missing (boundaryVal, or the centre value in the
// backward sweep :193-196,:279-282).
struct DirMask { bool sc, sw, se, sS, sN; };
__device__ __forceinline__ DirMask dir_mask(const YhK &k, const uint8_t *msk, int gi, int gj, int C) {
  DirMask d;
  if (k.solidSwitch) {
    d.sc = msk[C];
    d.sw = (gi > 0) && msk[C - 1];
    d.se = (gi < k.nx - 1) && msk[C + 1];
    const bool sn = (gj > 0) && msk[C - BP];           // mask at j-1
    const bool ss = (gj < k.nyg - 1) && msk[C + BP];   // mask at j+1
    d.sS = ss;   // the shipped code gates the (j-1) read by ss and the (j+1) read by sn (:182-183)
    d.sN = sn;
  } else {
    d.sc = true; d.sw = gi > 0; d.se = gi < (k.nx - 1); d.sS = gj > 0; d.sN = gj < (k.nyg - 1);
  }
  return d;
}

__global__ void __launch_bounds__(BTHREADS)
bfecc_kernel(const __grid_constant__ YhK k, const __grid_constant__ BfArgs a) {
  extern __shared__ __align__(16) double bsm[];
  double *g = bsm, *uf = bsm + BP * BP, *ue = bsm + 2 * BP * BP;
  double *sRx = bsm + 6 * BP * BP, *sRy = bsm + 7 * BP * BP;   // |c| dt / h, upwind direction in the sign bit
  uint8_t *msk = reinterpret_cast<uint8_t *>(bsm + 8 * BP * BP);
  // tile origin in GLOBAL rows: output rows are local rows [row0, row1) of a slab that holds global
  // rows [jg0, jg0 + ny) (whole sheet: jg0 = 0, row0 = 0, row1 = ny = nyg)
  const int ti0 = blockIdx.x * BT - BH, tj0 = k.jg0 + k.row0 + blockIdx.y * BT - BH;
  const int lrow_hi = k.jg0 + k.ny;   // global rows [jg0, lrow_hi) are held; the others read as 0.0
  const int tid = threadIdx.x;
  const bool neu = k.neumannBC != 0;
  const double tc = k.tc, bv = k.boundaryVal;
  CxyArgs cxy = a.cxy;
  if (a.sr_state) {
    cxy.cx = a.sr_state[YH_SR_C]; cxy.cy = a.sr_state[YH_SR_C + 1]; cxy.ct = a.sr_state[YH_SR_C + 2];
    cxy.cs = a.sr_state[YH_SR_CS]; cxy.sn = a.sr_state[YH_SR_SN];
  }

  // per-cell Courant numbers, once for the three sweeps and both fields (advFDBFECC.cu:30-34)
  for (int t = tid; t < BP * BP; t += BTHREADS) {
    const int gi = ti0 + t % BP, gj = tj0 + t / BP;
    double rx = 0.0, ry = 0.0;
    uint8_t m = 0;
    if (gi >= 0 && gi < k.nx && gj >= k.jg0 && gj < lrow_hi) {
      const int c = IDX(k, gi, gj);
      double ax, ay;
      if (a.from_c) {
        cxy_cell(k, cxy, gi, gj, ax, ay);
        const bool own = (t % BP >= BH) && (t % BP < BH + BT) && (t / BP >= BH) && (t / BP < BH + BT) &&
                         (gj < k.jg0 + k.row1);
        if (own && a.ax_out) { a.ax_out[c] = ax; a.ay_out[c] = ay; }
      } else { ax = a.ax[c]; ay = a.ay[c]; }
      const double cx = -ax, cy = -ay;
      const double Rx = sgn(cx) * cx * k.dt / k.hx, Ry = sgn(cy) * cy * k.dt / k.hy;
      rx = (cx > 0.0) ? fabs(Rx) : -fabs(Rx);   // Rx >= 0 always; sign bit carries (cx > 0)
      ry = (cy > 0.0) ? fabs(Ry) : -fabs(Ry);
      if (k.solidSwitch) m = a.solid[c];
    }
    sRx[t] = rx; sRy[t] = ry; msk[t] = m;
  }

  // Both fields go through every sweep together: the index selections, the upwind directions and
  // the Courant numbers of a cell are shared by u and v, and the pass has 4 barriers instead of 8.
  // Plane f of g / uf / ue is at offset f * FO.
  const double *gin[2] = {a.u_in, a.v_in};
  double *gout[2] = {a.u_out, a.v_out};
  __syncthreads();
  for (int t = tid; t < BP * BP; t += BTHREADS) {
    const int gi = ti0 + t % BP, gj = tj0 + t / BP;
    const bool in = gi >= 0 && gi < k.nx && gj >= k.jg0 && gj < lrow_hi;
    g[t] = in ? gin[0][IDX(k, gi, gj)] : 0.0;
    g[FO + t] = in ? gin[1][IDX(k, gi, gj)] : 0.0;
  }
  __syncthreads();
  // sweep 1 (forward) on the tile minus one ring
  for (int t = tid; t < BP * BP; t += BTHREADS) {
    const int li = t % BP, lj = t / BP, gi = ti0 + li, gj = tj0 + lj;
    if (li < 1 || li >= BP - 1 || lj < 1 || lj >= BP - 1) continue;
    if (gi < 0 || gi >= k.nx || gj < 0 || gj >= k.nyg) continue;
    const bool px = !signbit(sRx[t]), py = !signbit(sRy[t]);
    const double Rx = fabs(sRx[t]), Ry = fabs(sRy[t]);
    if (neu) {
      const NbrIdx x = neumann_idx(k, msk, gi, gj, ti0, tj0);
#pragma unroll
      for (int f = 0; f < 2; f++) {
        const double *gf = g + f * FO;
        const double FDx = px ? gf[t] - gf[x.W] : gf[t] - gf[x.E];
        const double FDy = py ? gf[t] - gf[x.S] : gf[t] - gf[x.N];
        uf[f * FO + t] = gf[t] - tc * (Rx * FDx + Ry * FDy);
      }
    } else {
      const DirMask d = dir_mask(k, msk, gi, gj, t);
      const int iS = TL(k, gi, gj - 1, ti0, tj0), iN = TL(k, gi, gj + 1, ti0, tj0);
#pragma unroll
      for (int f = 0; f < 2; f++) {
        const double *gf = g + f * FO;
        const double u = d.sc ? gf[t] : 0.0;
        const double W = d.sc && d.sw ? gf[t - 1] : (d.sc ? bv : 0.0);
        const double E = d.sc && d.se ? gf[t + 1] : (d.sc ? bv : 0.0);
        const double S = d.sc && d.sS ? gf[iS] : (d.sc ? bv : 0.0);
        const double N = d.sc && d.sN ? gf[iN] : (d.sc ? bv : 0.0);
        const double FDx = px ? u - W : u - E, FDy = py ? u - S : u - N;
        uf[f * FO + t] = u - tc * (Rx * FDx + Ry * FDy);
      }
    }
  }
  __syncthreads();
  // sweep 2 (backward + error compensation) on the tile minus two rings
  for (int t = tid; t < BP * BP; t += BTHREADS) {
    const int li = t % BP, lj = t / BP, gi = ti0 + li, gj = tj0 + lj;
    if (li < 2 || li >= BP - 2 || lj < 2 || lj >= BP - 2) continue;
    if (gi < 0 || gi >= k.nx || gj < 0 || gj >= k.nyg) continue;
    const bool px = !signbit(sRx[t]), py = !signbit(sRy[t]);
    const double Rx = fabs(sRx[t]), Ry = fabs(sRy[t]);
    if (neu) {
      const NbrIdx x = neumann_idx(k, msk, gi, gj, ti0, tj0);
#pragma unroll
      for (int f = 0; f < 2; f++) {
        const double *gf = g + f * FO, *uff = uf + f * FO;
        const double FDx = px ? uff[x.i2dE] - uff[x.E2] : uff[x.i2dW] - uff[x.W2];
        const double FDy = py ? uff[x.i2dN] - uff[x.N2] : uff[x.i2dS] - uff[x.S2];
        const double ub = uff[t] - tc * (Rx * FDx + Ry * FDy);
        ue[f * FO + t] = gf[t] - 0.5 * (ub - gf[t]);
      }
    } else {
      const DirMask d = dir_mask(k, msk, gi, gj, t);
      const int iS = TL(k, gi, gj - 1, ti0, tj0), iN = TL(k, gi, gj + 1, ti0, tj0);
#pragma unroll
      for (int f = 0; f < 2; f++) {
        const double *gf = g + f * FO, *uff = uf + f * FO;
        const double u = d.sc ? gf[t] : 0.0;
        const double uuf = d.sc ? uff[t] : 0.0;
        const double W = d.sc && d.sw ? uff[t - 1] : (d.sc ? uff[t] : 0.0);
        const double E = d.sc && d.se ? uff[t + 1] : (d.sc ? uff[t] : 0.0);
        const double S = d.sc && d.sS ? uff[iS] : (d.sc ? uff[t] : 0.0);
        const double N = d.sc && d.sN ? uff[iN] : (d.sc ? uff[t] : 0.0);
        const double FDx = px ? uuf - E : uuf - W, FDy = py ? uuf - N : uuf - S;
        // advFDBFECC.cu:287 ships "Rx*FDx - Ry*FDy" for u in the Dirichlet-square branch (B8)
        const double ub = (!k.solidSwitch && f == 0) ? uuf - tc * (Rx * FDx - Ry * FDy)
                                                     : uuf - tc * (Rx * FDx + Ry * FDy);
        ue[f * FO + t] = u - 0.5 * (ub - u);
      }
    }
  }
  __syncthreads();
  // sweep 3 (forward) on the 32 x 32 interior -> HBM
  for (int t = tid; t < BT * BT; t += BTHREADS) {
    const int li = BH + t % BT, lj = BH + t / BT, gi = ti0 + li, gj = tj0 + lj;
    if (gi >= k.nx || gj >= k.jg0 + k.row1) continue;
    const int q = li + BP * lj;
    const bool px = !signbit(sRx[q]), py = !signbit(sRy[q]);
    const double Rx = fabs(sRx[q]), Ry = fabs(sRy[q]);
    if (neu) {
      const NbrIdx x = neumann_idx(k, msk, gi, gj, ti0, tj0);
#pragma unroll
      for (int f = 0; f < 2; f++) {
        const double *uef = ue + f * FO;
        const double FDx = px ? uef[q] - uef[x.W] : uef[q] - uef[x.E];
        const double FDy = py ? uef[q] - uef[x.S] : uef[q] - uef[x.N];
        const double r = uef[q] - tc * (Rx * FDx + Ry * FDy);
        gout[f][IDX(k, gi, gj)] = x.sc ? r : 0.0;
      }
    } else {
      const DirMask d = dir_mask(k, msk, gi, gj, q);
      const int iS = TL(k, gi, gj - 1, ti0, tj0), iN = TL(k, gi, gj + 1, ti0, tj0);
#pragma unroll
      for (int f = 0; f < 2; f++) {
        const double *uef = ue + f * FO;
        const double uue = d.sc ? uef[q] : 0.0;
        const double W = d.sc && d.sw ? uef[q - 1] : (d.sc ? bv : 0.0);
        const double E = d.sc && d.se ? uef[q + 1] : (d.sc ? bv : 0.0);
        const double S = d.sc && d.sS ? uef[iS] : (d.sc ? bv : 0.0);
        const double N = d.sc && d.sN ? uef[iN] : (d.sc ? bv : 0.0);
        const double FDx = px ? uue - W : uue - E, FDy = py ? uue - S : uue - N;
        const double r = uue - tc * (Rx * FDx + Ry * FDy);
        gout[f][IDX(k, gi, gj)] = d.sc ? r : 0.0;
      }
    }
  }
}

// a whole sheet or a row slab of one (rows [jg0, jg0 + ny) of nx x ny_global, ghost rows included)
int check_slab(const yh_params *p) {
  YH_REQUIRE(p != nullptr, "null params");
  YH_REQUIRE(p->nx >= 8 && p->ny >= 1 && p->ny_global >= 8, "grid too small");
  YH_REQUIRE(p->jg0 >= 0 && p->jg0 + p->ny <= p->ny_global, "slab outside the global domain");
  return YH_OK;
}
// entry points whose result is a sum over the whole disc: a slab would silently drop rows
int check_sheet(const yh_params *p) {
  int rc = check_slab(p);
  if (rc != YH_OK) return rc;
  YH_REQUIRE(p->jg0 == 0 && p->ny_global == p->ny,
             "needs the whole sheet (row slabs: yh_sr_integral_rows + yh_sr_integrals_close)");
  return YH_OK;
}

}  // namespace

extern "C" {

int yh_slice(const yh_params *p, const double *u, const double *v, double *const slice[6],
             double *const slice0[6], int reduce_sym, int reduce_sym_start, const double *adv_x,
             const double *adv_y, int scheme, const int *tip_count, const yh_tip *tip_vector,
             int count, void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  if ((rc = check_slab(p)) != YH_OK) return rc;
  YH_REQUIRE(u && v && slice && adv_x && adv_y, "null pointer");
  YH_REQUIRE(scheme == 1 || scheme == 2, "scheme must be 1 or 2");
  YH_REQUIRE(!reduce_sym_start || slice0, "slice0 == NULL with reduce_sym_start");
  YH_REQUIRE(count == 0 || (tip_count && tip_vector), "tip list required when count != 0");
  if (!reduce_sym) return YH_OK;   // symmetryReduction.cu:127,:169: nothing is written
  YhK k = yh_make_k(p);
  SliceArgs a;
  a.u = u; a.v = v; a.ax = adv_x; a.ay = adv_y;
  for (int q = 0; q < 6; q++) {
    YH_REQUIRE(slice[q] != nullptr, "null slice array");
    a.s.p[q] = slice[q];
    a.s0.p[q] = reduce_sym_start ? slice0[q] : nullptr;
    YH_REQUIRE(!reduce_sym_start || slice0[q], "null slice0 array");
  }
  a.start = reduce_sym_start; a.scheme = scheme; a.count = count;
  a.tip_count = tip_count; a.tv = tip_vector; a.tipx0 = p->tipx0; a.tipy0 = p->tipy0;
  dim3 blk(32, 8), grd((p->nx + 31) / 32, (p->ny + 7) / 8);
  slice_kernel<<<grd, blk, 0, (cudaStream_t)stream>>>(k, a);
  YH_LAUNCH_CHECK();
  return YH_OK;
}

int yh_trapz(const yh_params *p, const double *const slice[6], const double *const slice0[6],
             const double *velTan_u, const double *velTan_v, double *integrals_host,
             const int *tip_count, const yh_tip *tip_vector, int count, void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  if ((rc = check_sheet(p)) != YH_OK) return rc;
  YH_REQUIRE(slice && slice0 && velTan_u && velTan_v && integrals_host, "null pointer");
  YH_REQUIRE(count == 0 || (tip_count && tip_vector), "tip list required when count != 0");
  IntArgs a;
  memset(&a, 0, sizeof(a));
  for (int q = 0; q < 6; q++) {
    YH_REQUIRE(slice[q] && slice0[q], "null slice array");
    a.s.p[q] = slice[q]; a.s0.p[q] = slice0[q];
  }
  a.vtu = velTan_u; a.vtv = velTan_v; a.count = count; a.tip_count = tip_count; a.tv = tip_vector;
  a.tipx0 = p->tipx0; a.tipy0 = p->tipy0;
  return run_integrals(p, a, false, integrals_host, (cudaStream_t)stream);
}

int yh_sr_integrals(const yh_params *p, const double *u, const double *v, const double *velTan_u,
                    const double *velTan_v, const double *adv_x, const double *adv_y,
                    double *integrals_host, const int *tip_count, const yh_tip *tip_vector,
                    int count, void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  if ((rc = check_sheet(p)) != YH_OK) return rc;
  YH_REQUIRE(u && v && velTan_u && velTan_v && adv_x && adv_y && integrals_host, "null pointer");
  YH_REQUIRE(count == 0 || (tip_count && tip_vector), "tip list required when count != 0");
  IntArgs a;
  memset(&a, 0, sizeof(a));
  a.u = u; a.v = v; a.ax = adv_x; a.ay = adv_y;
  a.vtu = velTan_u; a.vtv = velTan_v; a.count = count; a.tip_count = tip_count; a.tv = tip_vector;
  a.tipx0 = p->tipx0; a.tipy0 = p->tipy0;
  return run_integrals(p, a, true, integrals_host, (cudaStream_t)stream);
}

}  // extern "C"

int yh_sr_integrals_solve_device(const yh_params *p, const double *u, const double *v,
                                 const double *vtu, const double *vtv, const double *ax,
                                 const double *ay, const int *tip_count, const yh_tip *tv,
                                 double *sr_state, double *log_base, double step0, cudaStream_t st) {
  IntArgs a;
  memset(&a, 0, sizeof(a));
  a.u = u; a.v = v; a.ax = ax; a.ay = ay; a.vtu = vtu; a.vtv = vtv;
  a.count = 1;   // never the first step of a run: the disc follows the last tip
  a.tip_count = tip_count; a.tv = tv; a.tipx0 = p->tipx0; a.tipy0 = p->tipy0;
  a.sr = sr_state; a.log_base = log_base; a.step0 = step0; a.dt = p->dt;
  return run_integrals(p, a, true, nullptr, st);
}

int yh_sr_flush_phi_device(double *sr_state, double dt_phi, cudaStream_t st) {
  sr_flush_phi_kernel<<<1, 32, 0, st>>>(sr_state, dt_phi);
  YH_LAUNCH_CHECK();
  return YH_OK;
}

extern "C" {

static CxyArgs make_cxy(double *ax, double *ay, const uint8_t *solid, const double c[3], const double phi[3]) {
  CxyArgs a;
  a.ax = ax; a.ay = ay; a.solid = solid;
  a.cx = c[0]; a.cy = c[1]; a.ct = c[2];
  a.cs = cos(phi[2]); a.sn = sin(phi[2]);   // host libm, as solve_matrix (symmetryReduction.cu:390)
  return a;
}

int yh_cxy_field(const yh_params *p, double *adv_x, double *adv_y, const double c[3],
                 const double phi[3], const uint8_t *solid, void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  if ((rc = check_slab(p)) != YH_OK) return rc;
  YH_REQUIRE(adv_x && adv_y && c && phi, "null pointer");
  YH_REQUIRE(!p->solidSwitch || solid, "solidSwitch set but solid == NULL");
  YhK k = yh_make_k(p);
  CxyArgs a = make_cxy(adv_x, adv_y, solid, c, phi);
  dim3 blk(32, 8), grd((p->nx + 31) / 32, (p->ny + 7) / 8);
  cxy_kernel<<<grd, blk, 0, (cudaStream_t)stream>>>(k, a);
  YH_LAUNCH_CHECK();
  return YH_OK;
}

static int bfecc_common(const yh_params *p, BfArgs &a, void *stream, int row0 = 0, int row1 = -1) {
  YhK k = yh_make_k(p);
  if (row1 < 0) row1 = p->ny;
  YH_REQUIRE(row0 >= 0 && row0 < row1 && row1 <= p->ny, "bad row range");
  k.row0 = row0; k.row1 = row1;
  dim3 grd((p->nx + BT - 1) / BT, (row1 - row0 + BT - 1) / BT);
  const size_t smem = (size_t)8 * BP * BP * sizeof(double) + BP * BP;
  static bool attr_set[64] = {false};
  int dev = 0;
  YH_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    YH_CUDA(cudaFuncSetAttribute(bfecc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[dev & 63] = true;
  }
  bfecc_kernel<<<grd, BTHREADS, smem, (cudaStream_t)stream>>>(k, a);
  YH_LAUNCH_CHECK();
  return YH_OK;
}

int yh_advect_bfecc(const yh_params *p, const double *u_in, const double *v_in, double *u_out,
                    double *v_out, const double *adv_x, const double *adv_y, const uint8_t *solid,
                    void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  if ((rc = check_sheet(p)) != YH_OK) return rc;
  YH_REQUIRE(u_in && v_in && u_out && v_out && adv_x && adv_y, "null pointer");
  YH_REQUIRE(u_in != u_out && v_in != v_out, "in-place advection is not supported");
  YH_REQUIRE(!p->solidSwitch || solid, "solidSwitch set but solid == NULL");
  BfArgs a;
  memset(&a, 0, sizeof(a));
  a.u_in = u_in; a.v_in = v_in; a.u_out = u_out; a.v_out = v_out;
  a.ax = adv_x; a.ay = adv_y; a.solid = solid; a.from_c = 0;
  return bfecc_common(p, a, stream);
}

int yh_advect_bfecc_cphi(const yh_params *p, const double *u_in, const double *v_in,
                         double *u_out, double *v_out, const double c[3], const double phi[3],
                         double *adv_x, double *adv_y, const uint8_t *solid, void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  if ((rc = check_sheet(p)) != YH_OK) return rc;
  YH_REQUIRE(u_in && v_in && u_out && v_out && c && phi, "null pointer");
  YH_REQUIRE(u_in != u_out && v_in != v_out, "in-place advection is not supported");
  YH_REQUIRE((adv_x == nullptr) == (adv_y == nullptr), "adv_x / adv_y must both be set or NULL");
  YH_REQUIRE(!p->solidSwitch || solid, "solidSwitch set but solid == NULL");
  BfArgs a;
  memset(&a, 0, sizeof(a));
  a.u_in = u_in; a.v_in = v_in; a.u_out = u_out; a.v_out = v_out;
  a.ax_out = adv_x; a.ay_out = adv_y; a.solid = solid; a.from_c = 1;
  a.cxy = make_cxy(nullptr, nullptr, solid, c, phi);
  return bfecc_common(p, a, stream);
}

// Row-slab form (multi-GPU): rows [row0, row1) of the local arrays are written; they are
// valid when rows [row0-3, row1+3) (clipped to the global sheet) of u_in / v_in are held and valid.
int yh_advect_bfecc_cphi_rows(const yh_params *p, const double *u_in, const double *v_in,
                              double *u_out, double *v_out, const double c[3], const double phi[3],
                              double *adv_x, double *adv_y, const uint8_t *solid, int row0, int row1,
                              void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  if ((rc = check_slab(p)) != YH_OK) return rc;
  YH_REQUIRE(u_in && v_in && u_out && v_out && c && phi, "null pointer");
  YH_REQUIRE(u_in != u_out && v_in != v_out, "in-place advection is not supported");
  YH_REQUIRE((adv_x == nullptr) == (adv_y == nullptr), "adv_x / adv_y must both be set or NULL");
  YH_REQUIRE(!p->solidSwitch || solid, "solidSwitch set but solid == NULL");
  BfArgs a;
  memset(&a, 0, sizeof(a));
  a.u_in = u_in; a.v_in = v_in; a.u_out = u_out; a.v_out = v_out;
  a.ax_out = adv_x; a.ay_out = adv_y; a.solid = solid; a.from_c = 1;
  a.cxy = make_cxy(nullptr, nullptr, solid, c, phi);
  return bfecc_common(p, a, stream, row0, row1);
}

int yh_sr_disc_slots(const yh_params *p) {
  if (!p) return 0;
  return 2 * disc_radius_rows(p) + 1;
}

// The fused slice + trapz pass of a row slab: row sums of the 12 inner products for the disc rows
// this process owns (local rows [row0, row1)) -> rows_d[yh_sr_disc_slots(p) * 12] on the DEVICE;
// slots whose row belongs to another process are written as +0.0, so an element-wise sum over the
// processes (every slot has exactly one non-zero contributor: x + 0.0 == x) assembles the row sums of
// the whole sheet bit for bit.  (centre_x, centre_y): the disc centre every process agreed on -- the
// last tip of the gathered list, or (tipx0, tipy0).  No host sync.
int yh_sr_integral_rows(const yh_params *p, const double *u, const double *v, const double *velTan_u,
                        const double *velTan_v, const double *adv_x, const double *adv_y,
                        float centre_x, float centre_y, int row0, int row1, double *rows_d,
                        void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  if ((rc = check_slab(p)) != YH_OK) return rc;
  YH_REQUIRE(u && v && velTan_u && velTan_v && adv_x && adv_y && rows_d, "null pointer");
  YH_REQUIRE(row0 >= 0 && row0 <= row1 && row1 <= p->ny, "bad row range");
  // the tangent stencils reach two rows beyond a summed row
  const int lo = p->jg0 + row0 - 2, hi = p->jg0 + row1 + 2;
  YH_REQUIRE((lo < 0 || lo >= p->jg0) && (hi > p->ny_global || hi <= p->jg0 + p->ny),
             "summed rows need two ghost rows on each slab-internal side");
  IntArgs a;
  memset(&a, 0, sizeof(a));
  a.u = u; a.v = v; a.ax = adv_x; a.ay = adv_y; a.vtu = velTan_u; a.vtv = velTan_v;
  a.count = 0; a.tipx0 = centre_x; a.tipy0 = centre_y;   // disc_centre() returns exactly these
  a.rows = rows_d;                                        // a.done == NULL: no closing in this launch
  YhK k = yh_make_k(p);
  a.R = disc_radius_rows(p);
  a.own0 = p->jg0 + row0; a.own1 = p->jg0 + row1;
  integrals_rows_kernel<true><<<2 * a.R + 1, ROW_WARPS * 32, 0, (cudaStream_t)stream>>>(k, a);
  YH_LAUNCH_CHECK();
  return YH_OK;
}

// Totals of assembled row sums in the canonical order of the single-sheet pass -> integrals_host[12]
// (HOST, one sync), identical to yh_sr_integrals on the whole sheet.
int yh_sr_integrals_close(const yh_params *p, const double *rows_d, double *integrals_host, void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  if ((rc = check_slab(p)) != YH_OK) return rc;
  YH_REQUIRE(rows_d && integrals_host, "null pointer");
  double *out = nullptr;
  rc = yh_workspace(12 * sizeof(double), (void **)&out, 5);
  if (rc != YH_OK) return rc;
  integrals_close_kernel<<<1, 12 * 32, 0, (cudaStream_t)stream>>>(rows_d, 2 * disc_radius_rows(p) + 1, p->hx,
                                                                  p->hy, out);
  YH_LAUNCH_CHECK();
  return fetch12(out, integrals_host, (cudaStream_t)stream);
}

}  // extern "C"

int yh_advect_bfecc_device_c(const yh_params *p, const double *u_in, const double *v_in, double *u_out,
                             double *v_out, const double *sr_state, double *adv_x, double *adv_y,
                             const uint8_t *solid, cudaStream_t st) {
  BfArgs a;
  memset(&a, 0, sizeof(a));
  a.u_in = u_in; a.v_in = v_in; a.u_out = u_out; a.v_out = v_out;
  a.ax_out = adv_x; a.ay_out = adv_y; a.solid = solid; a.from_c = 1; a.sr_state = sr_state;
  a.cxy.solid = solid;
  return bfecc_common(p, a, (void *)st);
}
