// rd_fast.cu -- the HBM-roofline path of the monodomain step: Euler + 5-point Laplacian +
// no-flux (mirror) boundaries, square domain, with TEMPORAL BLOCKING (T time steps per pass
// over HBM).  Mode of reactionDiffusion.cu:186-199 + :498-513; same expression order, so the
// result is bit-identical to T launches of the reference kernel built with --fmad=false.
//
// Design ("3.5-D streaming"): a CTA owns a strip of W columns (W - 2H of them are outputs)
// and streams down RY rows.  Each time level l = 0..T-1 keeps a small ring of rows in shared
// memory; warp group l (W/2 threads, two cells per thread, 16-byte smem/global accesses)
// turns rows of level l-1 into one row of level l per iteration.  Level-0 rows arrive by
// cp.async (LDGSTS) 4 rows ahead of use; level-T rows go straight to HBM.  Every cell is read
// once and written once per T steps; the only redundancy is the 2H halo columns of a strip
// (3 % at W=256, T=4) and the 3T-row pipeline fill of a chunk.  One __syncthreads per row.
//
//   level-l row m is produced in iteration  m - c0 + 2l   (c0 = first level-0 row of the chunk)
//
// Zero-sign note: the reference forms the stage state as u0 + (0.0*0.0) (reactionDiffusion.cu:
// 117); x + 0.0 only changes -0.0 into +0.0.  Rows are stored canonicalised (+0.0) in the
// rings, which is exact for tc > 0 ("zero signs").
#include <stdlib.h>
#include <string.h>

#include "rd_euler_cell.cuh"

namespace {

using namespace yh_euler;

// T time levels, strip of W columns, one extra warp that only feeds level 0.
// The compute loop is unrolled by three so that the S / C / N row registers rotate by renaming
// instead of by moves; every sub-iteration ends in the CTA-wide barrier.
template <int T, int W, bool CANON, bool TC1, bool STIM, bool SOLID>
__global__ void __launch_bounds__(T *(W / 2) + 32)
rd_euler_stream(const __grid_constant__ YhK k, const __grid_constant__ FastArgs a) {
  constexpr int H = (T + 1) & ~1;        // halo columns each side (even: 16-byte alignment)
  constexpr int BX = W - 2 * H;          // output columns per strip
  constexpr int PITCH = W + 4;           // 2 pad doubles each side
  constexpr int ROW = 2 * PITCH;         // u plane then v plane
  constexpr int NR0 = 8, PF = 4, NRL = 4;
  constexpr int NTC = T * (W / 2);       // compute threads
  extern __shared__ __align__(16) double sm[];

  const int tid = threadIdx.x;
  const int nx = k.nx;
  const int x0 = blockIdx.x * BX, wx0 = x0 - H;
  const int y0 = k.row0 + blockIdx.y * a.RY;
  const int RYe = min(a.RY, k.row1 - y0);
  const int c0 = y0 - T;
  const int dom_lo = -k.jg0, dom_hi = k.nyg - k.jg0;   // local rows that exist globally
  const size_t zoff = (size_t)blockIdx.z * (size_t)a.sim_stride;
  const int n_it = ((RYe + 3 * T + 2) / 3) * 3;        // multiple of the unroll factor

  if (tid >= NTC) {
    // ---------------- loader warp: level-0 rows, PF rows ahead of their first use ----------
    const int lane = tid - NTC;
    constexpr int PER = W / 64;          // 16-byte pieces per lane per field per row
    const double *__restrict__ u_in = a.u_in + zoff;
    const double *__restrict__ v_in = a.v_in + zoff;
    const int ld_lo = max(dom_lo, c0), ld_hi = min(dom_hi, y0 + RYe + T);
    int goff[PER];
    bool ok[PER];
#pragma unroll
    for (int q = 0; q < PER; q++) {
      const int ggx = wx0 + 2 * (lane + 32 * q);
      ok[q] = (ggx >= 0) && (ggx < nx);
      goff[q] = ok[q] ? ggx : 0;
    }
    const unsigned sm0 = (unsigned)__cvta_generic_to_shared(sm) + (unsigned)(2 * lane + 2) * 8u;
    auto issue_row = [&](int q) {
      if (q >= ld_lo && q < ld_hi) {
        const unsigned dst = sm0 + (unsigned)((q - c0) & (NR0 - 1)) * (ROW * 8u);
        const double *ru = u_in + (size_t)q * nx;
        const double *rv = v_in + (size_t)q * nx;
#pragma unroll
        for (int p = 0; p < PER; p++) {
          cp_async16(dst + p * 512u, ru + goff[p], ok[p]);
          cp_async16(dst + PITCH * 8u + p * 512u, rv + goff[p], ok[p]);
        }
      }
      cp_async_commit();
    };
#pragma unroll
    for (int q = 0; q < PF; q++) issue_row(c0 + q);
    for (int it = 0; it < n_it; it++) {
      issue_row(c0 + it + PF);
      cp_async_wait<PF>();   // row c0 + it has landed
      __syncthreads();
    }
    cp_async_wait<0>();
    return;
  }

  // ---------------- compute warps: warp group `lev` turns level lev-1 rows into level lev ----
  const int lev = tid / (W / 2) + 1;     // warp-uniform
  const int c = 2 * (tid % (W / 2));     // window column of the thread's first cell
  const int gx = wx0 + c;
  const bool col_ok = (gx >= 0) && (gx < nx);
  const bool out_col = col_ok && (c >= H) && (c < W - H);
  const bool left_edge = (gx == 0), right_edge = (gx + 2 == nx);
  const bool canon = CANON && (lev == 1);   // level-0 data is raw: form u0 + (0.0*0.0) literally

  // rows this level must produce (empty for threads outside the domain)
  const int lo_l = max(dom_lo, y0 - (T - lev));
  const int hi_l = col_ok ? min(dom_hi, y0 + RYe + (T - lev)) : lo_l;

  // pacing (batched sweeps): level `lev` performs step count0 + lev - 1 of its sheet
  bool stim_on = false;
  if (STIM) {
    stim_on = k.stim != 0;
    if (a.period) {
      const int per = a.period[blockIdx.z];
      stim_on = per > 0 && ((a.count0 + lev - 1) % per) <= a.duration;
    }
  }

  const int src_mask = (lev == 1) ? (NR0 - 1) : (NRL - 1);
  const double *src_ring = sm + (lev == 1 ? 0 : (NR0 + (lev - 2) * NRL) * ROW) + c + 2;
  double *dst_ring = sm + (NR0 + (lev - 1) * NRL) * ROW + c + 2;
  double *gu = a.u_out + zoff + gx;
  double *gv = a.v_out + zoff + gx;
  const int m0 = c0 - 2 * lev;           // row handled in iteration 0

  auto ld_pair = [&](int row, double2 &pu, double2 &pv) {
    const double *r = src_ring + ((row - c0) & src_mask) * ROW;
    pu = *reinterpret_cast<const double2 *>(r);
    pv = *reinterpret_cast<const double2 *>(r + PITCH);
    if (canon) { pu.x += 0.0; pu.y += 0.0; pv.x += 0.0; pv.y += 0.0; }
  };

  // mask patterns of the pair, fetched two rows ahead of use (global / L2; 1 B per cell)
  unsigned pat_q0 = 0x1F1Fu, pat_q1 = 0x1F1Fu;
  auto ld_pat = [&](int row) -> unsigned {
    return *reinterpret_cast<const unsigned short *>(a.pat + (size_t)row * nx + gx);
  };

  // one row: S / C / N are the register-resident source rows m-1, m, m+1
  auto row_step = [&](int m, double2 &uS, double2 &vS, double2 &uC, double2 &vC, double2 &uN,
                      double2 &vN) {
    if (m >= lo_l && m < hi_l) {
      unsigned pat = 0x1F1Fu;
      if (SOLID) {
        if (m == lo_l) { pat_q0 = ld_pat(m); if (m + 1 < hi_l) pat_q1 = ld_pat(m + 1); }
        pat = pat_q0;
        pat_q0 = pat_q1;
        if (m + 2 < hi_l) pat_q1 = ld_pat(m + 2);
      }
      if (m == lo_l) {                   // first row of this level: nothing in registers yet
        ld_pair(m, uC, vC);
        ld_pair((m - 1 < dom_lo) ? m + 1 : m - 1, uS, vS);   // no-flux mirror at the first row
      }
      ld_pair((m + 1 >= dom_hi) ? m - 1 : m + 1, uN, vN);    // ... and at the last row
      const double *rc = src_ring + ((m - c0) & src_mask) * ROW;
      double uw = rc[-1], ue = rc[2], vw = rc[PITCH - 1], ve = rc[PITCH + 2];
      if (canon) { uw += 0.0; ue += 0.0; vw += 0.0; ve += 0.0; }
      if (left_edge) { uw = uC.y; vw = vC.y; }      // mirror: W of x=0 is x=1
      if (right_edge) { ue = uC.x; ve = vC.x; }     // mirror: E of x=nx-1 is x=nx-2
      bool s0 = false, s1 = false;
      if (STIM && stim_on) { s0 = yh_scs_on(k, gx, m + k.jg0); s1 = yh_scs_on(k, gx + 1, m + k.jg0); }
      double2 uo, vo;
      if (!SOLID || __all_sync(__activemask(), pat == 0x1F1Fu)) {   // all tissue around: plain stencil
        euler_cell<TC1>(k, uC.x, vC.x, uw, uC.y, uN.x, uS.x, vw, vC.y, vN.x, vS.x, s0, uo.x, vo.x);
        euler_cell<TC1>(k, uC.y, vC.y, uC.x, ue, uN.y, uS.y, vC.x, ve, vN.y, vS.y, s1, uo.y, vo.y);
      } else {
        euler_cell_solid<TC1>(k, pat & 0xFFu, uC.x, vC.x, uw, uC.y, uN.x, uS.x, vw, vC.y, vN.x, vS.x, s0, uo.x, vo.x);
        euler_cell_solid<TC1>(k, pat >> 8, uC.y, vC.y, uC.x, ue, uN.y, uS.y, vC.x, ve, vN.y, vS.y, s1, uo.y, vo.y);
      }
      if (STIM && a.apd.APD1 && out_col && m >= y0 && m < y0 + RYe) {   // fused sAPD epilogue (owner cells)
        const double th = 0.15;
        const bool e0 = ((uC.x > th) && (uo.x < th)) || ((uC.x < th) && (uo.x > th));
        const bool e1 = ((uC.y > th) && (uo.y < th)) || ((uC.y < th) && (uo.y > th));
        if (e0 || e1) {
          const size_t cidx = zoff + (size_t)m * nx + gx;
          if (e0) apd_event(k, a.apd, cidx, uo.x, uC.x, a.count0 + lev);
          if (e1) apd_event(k, a.apd, cidx + 1, uo.y, uC.y, a.count0 + lev);
        }
      }
      if (lev < T) {
        double *d = dst_ring + ((m - c0) & (NRL - 1)) * ROW;
        *reinterpret_cast<double2 *>(d) = uo;
        *reinterpret_cast<double2 *>(d + PITCH) = vo;
      } else if (out_col) {
        const size_t o = (size_t)m * nx;
        *reinterpret_cast<double2 *>(gu + o) = uo;
        *reinterpret_cast<double2 *>(gv + o) = vo;
      }
    }
    __syncthreads();
  };

  double2 uA = make_double2(0, 0), uB = uA, uC = uA, vA = uA, vB = uA, vC = uA;
  for (int it = 0; it < n_it; it += 3) {
    const int m = m0 + it;
    row_step(m, uA, vA, uB, vB, uC, vC);
    row_step(m + 1, uB, vB, uC, vC, uA, vA);
    row_step(m + 2, uC, vC, uA, vA, uB, vB);
  }
}

template <int T, int W, bool CANON, bool TC1, bool STIM, bool SOLID>
int launch3(const YhK &k, const FastArgs &a, int nsims, cudaStream_t st) {
  constexpr int H = (T + 1) & ~1, BX = W - 2 * H, PITCH = W + 4, ROW = 2 * PITCH;
  constexpr int NT = T * (W / 2) + 32;
  const size_t smem = (size_t)(8 + (T - 1) * 4) * ROW * sizeof(double);
  static bool attr_set[64] = {false};
  int dev = 0;
  YH_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    YH_CUDA(cudaFuncSetAttribute(rd_euler_stream<T, W, CANON, TC1, STIM, SOLID>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[dev & 63] = true;
  }
  static int slots[64] = {0};
  if (!slots[dev & 63]) {
    int per_sm = 1, sms = 148;
    YH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rd_euler_stream<T, W, CANON, TC1, STIM, SOLID>,
                                                          NT, smem));
    YH_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    slots[dev & 63] = (per_sm > 0 ? per_sm : 1) * sms;
  }
  const int rows = k.row1 - k.row0;
  FastArgs b = a;
  const int strips = (k.nx + BX - 1) / BX;
  b.RY = a.RY > 0 ? a.RY : pick_ry(rows, strips, nsims, T, slots[dev & 63]);
  dim3 grd(strips, (rows + b.RY - 1) / b.RY, nsims);
  auto kfn = rd_euler_stream<T, W, CANON, TC1, STIM, SOLID>;
  YH_LAUNCH(kfn, grd, NT, smem, st, k, b);
  return YH_OK;
}

template <int T, int W, bool CANON, bool TC1>
int launch2(const YhK &k, const FastArgs &a, int nsims, cudaStream_t st) {
  const bool stim = k.stim != 0 || a.period != nullptr || a.apd.APD1 != nullptr;
  if (a.pat)   // obstacle masks: one variant (STIM on) keeps the instantiation count down
    return launch3<T, W, CANON, TC1, true, true>(k, a, nsims, st);
  return stim ? launch3<T, W, CANON, TC1, true, false>(k, a, nsims, st)
              : launch3<T, W, CANON, TC1, false, false>(k, a, nsims, st);
}

template <int T, int W>
int launch(const YhK &k, const FastArgs &a, int nsims, bool canon, cudaStream_t st) {
  // DEF variant: every constant whose operation is an exact identity at the reference defaults
  const bool tc1 = (k.tc == 1.0) && (k.mu == 1.0) && (k.delta == 1.0) && (k.gamma == 0.0) && (k.theta == 0.0) &&
                   (k.gateDiff != 0);   // the reference's default constants AND switches
  if (canon) return tc1 ? launch2<T, W, true, true>(k, a, nsims, st) : launch2<T, W, true, false>(k, a, nsims, st);
  return tc1 ? launch2<T, W, false, true>(k, a, nsims, st) : launch2<T, W, false, false>(k, a, nsims, st);
}

// pat[c] = sc | sw<<1 | se<<2 | sn<<3 | ss<<4 with the neighbour indices of
// reactionDiffusion.cu:149-152 (mirrored at the GLOBAL edges); rows whose vertical neighbour is
// not stored locally (outermost ghost rows of a slab) are never computed, they get a clamped index.
__global__ void solid_pattern_kernel(const __grid_constant__ YhK k, const uint8_t *__restrict__ solid,
                                     uint8_t *__restrict__ pat) {
  const long long n = (long long)k.nx * k.ny;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n;
       c += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(c % k.nx), j = (int)(c / k.nx), gj = j + k.jg0;
    int jS = yh_mir(gj - 1, k.nyg) - k.jg0, jN = yh_mir(gj + 1, k.nyg) - k.jg0;
    jS = min(max(jS, 0), k.ny - 1); jN = min(max(jN, 0), k.ny - 1);
    const unsigned sc = solid[c] != 0;
    const unsigned sw = solid[(size_t)j * k.nx + yh_mir(i - 1, k.nx)] != 0;
    const unsigned se = solid[(size_t)j * k.nx + yh_mir(i + 1, k.nx)] != 0;
    const unsigned sn = solid[(size_t)jN * k.nx + i] != 0;
    const unsigned ss = solid[(size_t)jS * k.nx + i] != 0;
    pat[c] = (uint8_t)(sc | (sw << 1) | (se << 2) | (sn << 3) | (ss << 4));
  }
}

}  // namespace

// Euler / 5-point / no-flux; the obstacle-mask mode is covered too when the caller supplies the
// pattern array (yh_rd_solid_patterns), see yh_rd_fast_solid_supported.
static int fast_mode_ok(const YhK &k, int tb) {
  if (k.timeIntOrder != 1 || !k.neumannBC) return 0;
  if (!(k.tc > 0.0)) return 0;            // zero-sign argument needs tc > 0
  if ((k.nx & 1) || k.nx < 8) return 0;   // two cells per thread, 16-byte rows
  if (tb != 1 && tb != 2 && tb != 4) return 0;
  return 1;
}
int yh_rd_fast_supported(const YhK &k, int tb) {
  if (k.lap4 || k.solidSwitch || k.anisotropy) return 0;
  return fast_mode_ok(k, tb);
}
// lap4 and anisotropy are ignored in the mask branch (reactionDiffusion.cu:154-184)
int yh_rd_fast_solid_supported(const YhK &k, int tb) {
  return k.solidSwitch ? fast_mode_ok(k, tb) : 0;
}

int yh_rd_solid_patterns(const YhK &k, const uint8_t *solid, uint8_t *pat, cudaStream_t st) {
  const long long n = (long long)k.nx * k.ny;
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  YH_LAUNCH(solid_pattern_kernel, blocks, 256, 0, st, k, solid, pat);
  return YH_OK;
}

int yh_launch_rd_quad_paced(const YhK &k, int tb, const yh_euler::FastArgs &a, int nsims, bool canon, int W,
                            cudaStream_t st);   // rd_quad.cu

int yh_launch_rd_fast_paced(const YhK &k, int tb, const double *u_in, const double *v_in,
                            double *u_out, double *v_out, int nsims, long long sim_stride,
                            const int *period_d, int duration_it, int count0, int canon_in,
                            cudaStream_t st, const YhApd *apd, const uint8_t *pat) {
  if (!(pat ? yh_rd_fast_solid_supported(k, tb) : yh_rd_fast_supported(k, tb))) return YH_ERR_UNSUPPORTED;
  FastArgs a{u_in, v_in, u_out, v_out, 0, sim_stride, period_d, duration_it, count0, {}, pat, 0, 0, 0, 0, 0, 0, 0, 0};
  if (apd) a.apd = *apd; else memset(&a.apd, 0, sizeof(a.apd));
  const int rows = k.row1 - k.row0;
  if (rows <= 0) return YH_OK;
  const char *force_w = getenv("YH_FAST_W");   // tuning hook
  const char *force_ry = getenv("YH_FAST_RY");  // tuning hook
  a.RY = force_ry ? atoi(force_ry) : 0;                // 0: chosen per kernel variant
  const bool canon = canon_in != 0;
  // strip width: narrow strips when the sheet is too small to fill 148 SMs otherwise (latency-bound
  // regime).  For big sheets W = 128 beats W = 256 (16384^2, T = 4: 331 vs 316 Gcell/s) although it
  // doubles the halo redundancy (6 % vs 3 %): four resident CTAs of 9 warps are four independent
  // per-row barrier domains instead of two of 17 warps.
  const long long cells = (long long)k.nx * rows * nsims;
  int W = cells >= (1ll << 21) ? 128 : 64;
  if (force_w) W = atoi(force_w);
  // four columns per thread (rd_quad.cu) for sheets that fill the machine; YH_EULER_KERNEL = quad | pair overrides
  const char *kern = getenv("YH_EULER_KERNEL");
  // measured (B200, 16384^2 / 8192^2, Gcell/s): T = 4 quad 410 vs pair 361; T = 2 332 vs 346; T = 1 185 (TMA feed) vs 160
  // batched paced sheets with the fused APD epilogue (C5): pair 167 vs quad 153 -- stays on the pair kernel
  const bool paced = nsims > 1 || period_d != nullptr || apd != nullptr;
  const bool quad = kern ? (kern[0] == 'q') : (tb != 2 && !paced && cells >= (1ll << 21));
  if (quad && k.nx >= 16) return yh_launch_rd_quad_paced(k, tb, a, nsims, canon, W == 256 ? 256 : 128, st);
#define YH_FAST_DISPATCH(WW)                                     \
  switch (tb) {                                                  \
    case 1: return launch<1, WW>(k, a, nsims, canon, st);        \
    case 2: return launch<2, WW>(k, a, nsims, canon, st);        \
    default: return launch<4, WW>(k, a, nsims, canon, st);       \
  }
  if (W == 64) { YH_FAST_DISPATCH(64) }
  if (W == 128) { YH_FAST_DISPATCH(128) }
  YH_FAST_DISPATCH(256)
#undef YH_FAST_DISPATCH
}

int yh_launch_rd_fast(const YhK &k, int tb, const double *u_in, const double *v_in,
                      double *u_out, double *v_out, const uint8_t *pat, int canon_in,
                      cudaStream_t st) {
  return yh_launch_rd_fast_paced(k, tb, u_in, v_in, u_out, v_out, 1, 0, nullptr, 0, 0, canon_in, st,
                                 nullptr, pat);
}
