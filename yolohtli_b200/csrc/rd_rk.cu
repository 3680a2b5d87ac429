// rd_rk.cu -- the reference's DEFAULT mode in one launch per time step: Euler / RK2 / RK4 with
// SYNCHRONOUS stages, the optional 4th-order (9-point + J correction) Laplacian and the obstacle
// masks (reactionDiffusion.cu:71-93, 115-247, 154-184, 498-561), no-flux boundaries.
//
// Same 3.5-D streaming skeleton as rd_fast.cu, with the RK stages as pipeline levels: a CTA
// owns a strip of W columns and streams down the rows; warp group P canonicalises a level-0 row
// and evaluates the ionic currents J = (I_sum, I_v) once per cell; warp group S_k turns three
// rows of stage-k state (U, V, J) into du_k, accumulates rhs += w_k du_k, and emits the
// stage-(k+1) row  U = u0 + a_{k+1} du_k, its J, and the running rhs -- all through small
// shared-memory rings, one CTA barrier per row.  Every cell is read once and written once per
// time step (plus velTan when asked), no J / stage arrays in HBM, no in-place races.
//
//   row m:  P at iteration m-c0+1,  S_k at iteration m-c0+1+2(k+1)
//
// This mode is FP64-pipe bound (~350 DADD/DMUL per cell-step), not HBM bound.
#include <stdlib.h>

#include "yh_common.cuh"

namespace {

struct RkArgs {
  const double *u_in, *v_in;
  double *u_out, *v_out, *vtu, *vtv;
  const uint8_t *solid;   // SOLID variants: 1 = tissue (main.cu:676-680), local rows
  int RY;
};

// Obstacle masks (reactionDiffusion.cu:154-184): the six stencil coefficients of a cell depend
// only on the mask of the cell and its four (mirrored) neighbours, not on the stage, so group P
// packs them once per cell into 13 bits -- cxx cxy cxz cyx cyy cyz (2 bits each, values 0/1/2)
// and sc -- and the stage groups read the code of THEIR cell only: no mask halo in the pipeline.
__device__ __forceinline__ unsigned solid_code(bool sc, bool sw, bool se, bool sn, bool ss) {
  const unsigned cxx = (sw && se) && (sw && sc) ? 1u : ((sw && sc) ? 2u : 0u);
  const unsigned cxy = sc ? ((sw || se) ? 2u : 0u) : 0u;
  const unsigned cxz = (sw && se) && (sc && se) ? 1u : ((sc && se) ? 2u : 0u);
  const unsigned cyx = (sn && ss) && (sn && sc) ? 1u : ((sn && sc) ? 2u : 0u);
  const unsigned cyy = sc ? ((sn || ss) ? 2u : 0u) : 0u;
  const unsigned cyz = (sn && ss) && (sc && ss) ? 1u : ((sc && ss) ? 2u : 0u);
  return cxx | (cxy << 2) | (cxz << 4) | (cyx << 6) | (cyy << 8) | (cyz << 10) | ((sc ? 1u : 0u) << 12);
}
// 0 / 1 / 2 as an exact double without a conversion instruction
__device__ __forceinline__ double coef(unsigned code, int shift) {
  const unsigned c2 = (code >> shift) & 3u;
  return __hiloint2double(c2 ? (int)(0x3FE00000u + (c2 << 20)) : 0, 0);
}

__device__ __forceinline__ void cp_async16(unsigned smem, const void *gmem, bool valid) {
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem), "l"(gmem), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// K stages, strip of W columns.  Arrays of one ring row: U V Ju Jv ru rv (6 x PITCH doubles).
template <int K, int W, bool LAP4, bool SOLID>
__global__ void __launch_bounds__((K + 1) * (W / 2) + 32)
rd_rk_stream(const __grid_constant__ YhK k, const __grid_constant__ RkArgs a) {
  constexpr int H = (K + 1) & ~1;        // halo columns each side (even: 16-byte alignment)
  constexpr int BX = W - 2 * H;
  constexpr int PITCH = W + 4;
  constexpr int ROW0 = 2 * PITCH;        // level-0 ring row: u0 v0
  constexpr int ROWA = 6 * PITCH;        // stage ring row
  constexpr int NR0 = 16, PF = 4, NRA = 4;
  constexpr int NTC = (K + 1) * (W / 2);
  extern __shared__ __align__(16) double sm[];
  double *R0 = sm;
  double *A = sm + NR0 * ROW0;           // A[kk] = A + kk * NRA * ROWA
  unsigned short *Cr = reinterpret_cast<unsigned short *>(A + K * NRA * ROWA);   // [NR0][W] mask codes

  const int tid = threadIdx.x;
  const int nx = k.nx;
  const int x0 = blockIdx.x * BX, wx0 = x0 - H;
  const int y0 = k.row0 + blockIdx.y * a.RY;
  const int RYe = min(a.RY, k.row1 - y0);
  const int c0 = y0 - K;
  const int dom_lo = -k.jg0, dom_hi = k.nyg - k.jg0;
  const int n_it = RYe + 3 * K + 1;

  if (tid >= NTC) {   // ---------------- loader warp ----------------
    const int lane = tid - NTC;
    constexpr int PER = (W / 2 + 31) / 32;
    const int ld_lo = max(dom_lo, c0), ld_hi = min(dom_hi, y0 + RYe + K);
    int goff[PER];
    bool ok[PER], use[PER];
#pragma unroll
    for (int q = 0; q < PER; q++) {
      const int pr = lane + 32 * q;
      const int ggx = wx0 + 2 * pr;
      use[q] = pr < W / 2;
      ok[q] = use[q] && (ggx >= 0) && (ggx < nx);
      goff[q] = ok[q] ? ggx : 0;
    }
    const unsigned sm0 = (unsigned)__cvta_generic_to_shared(R0) + (unsigned)(2 * lane + 2) * 8u;
    auto issue_row = [&](int q) {
      if (q >= ld_lo && q < ld_hi) {
        const unsigned dst = sm0 + (unsigned)((q - c0) & (NR0 - 1)) * (ROW0 * 8u);
        const double *ru = a.u_in + (size_t)q * nx;
        const double *rv = a.v_in + (size_t)q * nx;
#pragma unroll
        for (int p = 0; p < PER; p++) {
          if (use[p]) {
            cp_async16(dst + p * 512u, ru + goff[p], ok[p]);
            cp_async16(dst + PITCH * 8u + p * 512u, rv + goff[p], ok[p]);
          }
        }
      }
      cp_async_commit();
    };
#pragma unroll
    for (int q = 0; q < PF; q++) issue_row(c0 + q);
    for (int it = 0; it < n_it; it++) {
      issue_row(c0 + it + PF);
      cp_async_wait<PF>();
      __syncthreads();
    }
    cp_async_wait<0>();
    return;
  }

  const int g = tid / (W / 2);           // 0 = P, 1..K = stage g-1 (warp-uniform)
  const int c = 2 * (tid % (W / 2));
  const int gx = wx0 + c;
  const bool col_ok = (gx >= 0) && (gx < nx);
  const bool out_col = col_ok && (c >= H) && (c < W - H);
  const bool left_edge = (gx == 0), right_edge = (gx + 2 == nx);
  const int cc = c + 2;                  // column offset inside a padded ring row

  // Runge-Kutta tables (reactionDiffusion.cu:71-86)
  double a_next = 0.0, w_k = 0.0;
  const int st = g - 1;                  // stage index of this group
  if (K == 4) {
    const double ki[5] = {0.0, 0.5, 0.5, 1.0, 0.0};
    const double ws[4] = {0.166666666666667, 0.333333333333333, 0.333333333333333, 0.166666666666667};
    if (st >= 0) { a_next = ki[st + 1]; w_k = ws[st]; }
  } else if (K == 2) {
    if (st == 0) { a_next = 0.5; w_k = 0.0; }
    if (st == 1) { a_next = 0.0; w_k = 1.0; }
  } else {   // Euler
    w_k = 1.0;
  }

  // rows each group handles (empty for out-of-domain columns)
  int lo_g, hi_g;
  if (g == 0) { lo_g = max(dom_lo, y0 - K); hi_g = min(dom_hi, y0 + RYe + K); }
  else { lo_g = max(dom_lo, y0 - (K - 1 - st)); hi_g = min(dom_hi, y0 + RYe + (K - 1 - st)); }
  if (!col_ok) hi_g = lo_g;
  const int m_shift = (g == 0) ? 1 : 1 + 2 * g;   // row m = it + c0 - m_shift

  const double q4 = k.qx4 + k.qy4;
  const double m2q = -2.0 * q4;                    // -2.0*( qx4+qy4 )
  const double mrs2 = -k.rscale * 2.0;             // -rscale*2.0  (then *q4, :235)
  const double rsq = k.rscale * q4;                // rscale*( qx4+qy4 ) (:239)

  for (int it = 0; it < n_it; it++) {
    const int m = it + c0 - m_shift;
    if (m >= lo_g && m < hi_g) {
      const int slot = (m - c0) & (NRA - 1);
      const int gj = m + k.jg0;
      if (g == 0) {
        // ---- P: stage-0 state u0 + (0.0*0.0) and its currents ----
        const double *r0 = R0 + ((m - c0) & (NR0 - 1)) * ROW0 + cc;
        double2 u = *reinterpret_cast<const double2 *>(r0);
        double2 v = *reinterpret_cast<const double2 *>(r0 + PITCH);
        u.x += 0.0; u.y += 0.0; v.x += 0.0; v.y += 0.0;
        double2 ju, jv;
        ju.x = yh_Isum(k, u.x, v.x, yh_scs(k, gx, gj));
        ju.y = yh_Isum(k, u.y, v.y, yh_scs(k, gx + 1, gj));
        jv.x = yh_Iv(k, u.x, v.x);
        jv.y = yh_Iv(k, u.y, v.y);
        double *d = A + slot * ROWA + cc;
        *reinterpret_cast<double2 *>(d) = u;
        *reinterpret_cast<double2 *>(d + PITCH) = v;
        *reinterpret_cast<double2 *>(d + 2 * PITCH) = ju;
        *reinterpret_cast<double2 *>(d + 3 * PITCH) = jv;
        if (SOLID) {
          const int ny = k.nyg;
          const uint8_t *mc = a.solid + (size_t)m * nx;
          const uint8_t *mS = a.solid + (size_t)(yh_mir(gj - 1, ny) - k.jg0) * nx;   // S = j-1 (:149)
          const uint8_t *mN = a.solid + (size_t)(yh_mir(gj + 1, ny) - k.jg0) * nx;   // N = j+1 (:150)
          const bool c0m = mc[gx] != 0, c1m = mc[gx + 1] != 0;
          const bool w0 = mc[yh_mir(gx - 1, nx)] != 0, e1 = mc[yh_mir(gx + 2, nx)] != 0;
          const unsigned code0 = solid_code(c0m, w0, c1m, mN[gx] != 0, mS[gx] != 0);
          const unsigned code1 = solid_code(c1m, c0m, e1, mN[gx + 1] != 0, mS[gx + 1] != 0);
          *reinterpret_cast<unsigned *>(Cr + ((m - c0) & (NR0 - 1)) * W + c) = code0 | (code1 << 16);
        }
      } else {
        // ---- S_st: du of stage st for the pair (c, c+1) of row m ----
        const double *Ak = A + st * NRA * ROWA;
        const int ms = (m - 1 < dom_lo) ? m + 1 : m - 1;
        const int mn = (m + 1 >= dom_hi) ? m - 1 : m + 1;
        const double *rc = Ak + slot * ROWA + cc;
        const double *rs = Ak + ((ms - c0) & (NRA - 1)) * ROWA + cc;
        const double *rn = Ak + ((mn - c0) & (NRA - 1)) * ROWA + cc;
        double du[2], dv[2];
        unsigned codes = 0;
        if (SOLID) codes = *reinterpret_cast<const unsigned *>(Cr + ((m - c0) & (NR0 - 1)) * W + c);
#pragma unroll
        for (int f = 0; f < 2; f++) {   // f = 0: u with Ju, f = 1: v with Jv
          const double *pc = rc + f * PITCH, *ps = rs + f * PITCH, *pn = rn + f * PITCH;
          const double2 C = *reinterpret_cast<const double2 *>(pc);
          const double2 S = *reinterpret_cast<const double2 *>(ps);
          const double2 N = *reinterpret_cast<const double2 *>(pn);
          double Wv = pc[-1], Ev = pc[2];
          if (left_edge) Wv = C.y;
          if (right_edge) Ev = C.x;
          double d0, d1;
          if (SOLID) {   // reactionDiffusion.cu:171-180
            const unsigned k0 = codes & 0xFFFFu, k1 = codes >> 16;
            if (f == 0) {
              d0 = ((coef(k0, 0) * Wv - coef(k0, 2) * C.x + coef(k0, 4) * C.y) * k.rx +
                    (coef(k0, 6) * N.x - coef(k0, 8) * C.x + coef(k0, 10) * S.x) * k.ry);
              d1 = ((coef(k1, 0) * C.x - coef(k1, 2) * C.y + coef(k1, 4) * Ev) * k.rx +
                    (coef(k1, 6) * N.y - coef(k1, 8) * C.y + coef(k1, 10) * S.y) * k.ry);
            } else if (k.gateDiff) {
              d0 = ((coef(k0, 0) * Wv - coef(k0, 2) * C.x + coef(k0, 4) * C.y) * k.rx * k.rscale +
                    (coef(k0, 6) * N.x - coef(k0, 8) * C.x + coef(k0, 10) * S.x) * k.ry * k.rscale);
              d1 = ((coef(k1, 0) * C.x - coef(k1, 2) * C.y + coef(k1, 4) * Ev) * k.rx * k.rscale +
                    (coef(k1, 6) * N.y - coef(k1, 8) * C.y + coef(k1, 10) * S.y) * k.ry * k.rscale);
            } else {
              d0 = 0.0; d1 = 0.0;
            }
          } else if (f == 0) {
            d0 = ((Wv - 2.0 * C.x + C.y) * k.rx + (N.x - 2.0 * C.x + S.x) * k.ry);
            d1 = ((C.x - 2.0 * C.y + Ev) * k.rx + (N.y - 2.0 * C.y + S.y) * k.ry);
          } else if (k.gateDiff) {
            d0 = ((Wv - 2.0 * C.x + C.y) * k.rx * k.rscale + (N.x - 2.0 * C.x + S.x) * k.ry * k.rscale);
            d1 = ((C.x - 2.0 * C.y + Ev) * k.rx * k.rscale + (N.y - 2.0 * C.y + S.y) * k.ry * k.rscale);
          } else {
            d0 = 0.0; d1 = 0.0;
          }
          const double2 Jc = *reinterpret_cast<const double2 *>(pc + 2 * PITCH);
          if (LAP4 && (f == 0 || k.gateDiff)) {
            double SWv = ps[-1], SEv = ps[2], NWv = pn[-1], NEv = pn[2];
            if (left_edge) { SWv = S.y; NWv = N.y; }
            if (right_edge) { SEv = S.x; NEv = N.x; }
            const double2 Js = *reinterpret_cast<const double2 *>(ps + 2 * PITCH);
            const double2 Jn = *reinterpret_cast<const double2 *>(pn + 2 * PITCH);
            double JW = pc[2 * PITCH - 1], JE = pc[2 * PITCH + 2];
            if (left_edge) JW = Jc.y;
            if (right_edge) JE = Jc.x;
            if (f == 0) {   // reactionDiffusion.cu:221-229
              d0 += m2q * (+(Wv - C.x + C.y) + (N.x - C.x + S.x));
              d1 += m2q * (+(C.x - C.y + Ev) + (N.y - C.y + S.y));
              d0 += q4 * (SWv + S.y + NWv + N.y);
              d1 += q4 * (S.x + SEv + N.x + NEv);
            } else {        // :235-239
              d0 += mrs2 * q4 * (+(Wv - C.x + C.y) + (N.x - C.x + S.x));
              d1 += mrs2 * q4 * (+(C.x - C.y + Ev) + (N.y - C.y + S.y));
              d0 += rsq * (SWv + S.y + NWv + N.y);
              d1 += rsq * (S.x + SEv + N.x + NEv);
            }
            d0 -= ((JW - 2.0 * Jc.x + Jc.y) * k.fx4 + (Jn.x - 2.0 * Jc.x + Js.x) * k.fy4);
            d1 -= ((Jc.x - 2.0 * Jc.y + JE) * k.fx4 + (Jn.y - 2.0 * Jc.y + Js.y) * k.fy4);
          }
          d0 -= k.dt * Jc.x;   // :498-499
          d1 -= k.dt * Jc.y;
          if (f == 0) { du[0] = d0; du[1] = d1; }
          else { dv[0] = d0; dv[1] = d1; }
        }
        // running rhs (:502-503)
        double2 ru = make_double2(0.0, 0.0), rv = ru;
        if (st > 0) {
          ru = *reinterpret_cast<const double2 *>(rc + 4 * PITCH);
          rv = *reinterpret_cast<const double2 *>(rc + 5 * PITCH);
        }
        ru.x += (w_k * du[0]); ru.y += (w_k * du[1]);
        rv.x += (w_k * dv[0]); rv.y += (w_k * dv[1]);
        const double *r0 = R0 + ((m - c0) & (NR0 - 1)) * ROW0 + cc;
        const double2 u0 = *reinterpret_cast<const double2 *>(r0);
        const double2 v0 = *reinterpret_cast<const double2 *>(r0 + PITCH);
        if (st < K - 1) {
          // stage st+1 state (:117-118), its currents, and the rhs travel to the next group
          double2 un, vn, ju, jv;
          un.x = u0.x + (a_next * du[0]); un.y = u0.y + (a_next * du[1]);
          vn.x = v0.x + (a_next * dv[0]); vn.y = v0.y + (a_next * dv[1]);
          ju.x = yh_Isum(k, un.x, vn.x, yh_scs(k, gx, gj));
          ju.y = yh_Isum(k, un.y, vn.y, yh_scs(k, gx + 1, gj));
          jv.x = yh_Iv(k, un.x, vn.x);
          jv.y = yh_Iv(k, un.y, vn.y);
          double *d = A + (st + 1) * NRA * ROWA + slot * ROWA + cc;
          *reinterpret_cast<double2 *>(d) = un;
          *reinterpret_cast<double2 *>(d + PITCH) = vn;
          *reinterpret_cast<double2 *>(d + 2 * PITCH) = ju;
          *reinterpret_cast<double2 *>(d + 3 * PITCH) = jv;
          *reinterpret_cast<double2 *>(d + 4 * PITCH) = ru;
          *reinterpret_cast<double2 *>(d + 5 * PITCH) = rv;
        } else if (out_col) {
          double2 uo, vo;   // :512-513
          uo.x = u0.x + k.tc * ru.x; uo.y = u0.y + k.tc * ru.y;
          vo.x = v0.x + k.tc * rv.x; vo.y = v0.y + k.tc * rv.y;
          const bool sc0 = !SOLID || ((codes >> 12) & 1u), sc1 = !SOLID || ((codes >> 28) & 1u);
          if (SOLID) {   // :521-522 masked cells are exactly 0.0
            uo.x = sc0 ? uo.x : 0.0; uo.y = sc1 ? uo.y : 0.0;
            vo.x = sc0 ? vo.x : 0.0; vo.y = sc1 ? vo.y : 0.0;
          }
          const size_t o = (size_t)m * nx + gx;
          *reinterpret_cast<double2 *>(a.u_out + o) = uo;
          *reinterpret_cast<double2 *>(a.v_out + o) = vo;
          if (a.vtu && k.gateDiff) {   // :551-552
            double2 tu, tv;
            tu.x = sc0 ? ru.x / k.dt : 0.0; tu.y = sc1 ? ru.y / k.dt : 0.0;   // :529-530
            tv.x = sc0 ? rv.x / k.dt : 0.0; tv.y = sc1 ? rv.y / k.dt : 0.0;
            *reinterpret_cast<double2 *>(a.vtu + o) = tu;
            *reinterpret_cast<double2 *>(a.vtv + o) = tv;
          }
        }
      }
    }
    __syncthreads();
  }
}

static int pick_ry_rk(int rows, int strips, int K, int slots) {
  const int min_ry = 4;
  int best_ry = rows;
  double best = -1.0;
  const int cmax = rows / min_ry > 0 ? rows / min_ry : 1;
  for (int C = 1; C <= cmax; C++) {
    const int ry = (rows + C - 1) / C;
    const int chunks = (rows + ry - 1) / ry;
    const long long ctas = (long long)strips * chunks;
    const long long waves = (ctas + slots - 1) / slots;
    const double eff = (double)ctas / (double)(waves * slots) * (double)ry / (double)(ry + 3 * K + 1);
    if (eff > best + 1e-9) { best = eff; best_ry = ry; }
    if (ry <= min_ry) break;
  }
  return best_ry;
}

template <int K, int W, bool LAP4, bool SOLID>
int launch(const YhK &k, RkArgs a, cudaStream_t st) {
  constexpr int PITCH = W + 4, BX = W - 2 * ((K + 1) & ~1);
  constexpr int NT = (K + 1) * (W / 2) + 32;
  const size_t smem = ((size_t)16 * 2 * PITCH + (size_t)K * 4 * 6 * PITCH) * sizeof(double) +
                      (size_t)16 * W * sizeof(unsigned short);
  static bool attr_set[64] = {false};
  static int slots[64] = {0};
  int dev = 0;
  YH_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    YH_CUDA(cudaFuncSetAttribute(rd_rk_stream<K, W, LAP4, SOLID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1, sms = 148;
    YH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rd_rk_stream<K, W, LAP4, SOLID>, NT, smem));
    YH_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    slots[dev & 63] = (per_sm > 0 ? per_sm : 1) * sms;
    attr_set[dev & 63] = true;
  }
  const int rows = k.row1 - k.row0;
  const int strips = (k.nx + BX - 1) / BX;
  const char *force_ry = getenv("YH_RK_RY");
  a.RY = force_ry ? atoi(force_ry) : pick_ry_rk(rows, strips, K, slots[dev & 63]);
  dim3 grd(strips, (rows + a.RY - 1) / a.RY);
  rd_rk_stream<K, W, LAP4, SOLID><<<grd, NT, smem, st>>>(k, a);
  YH_LAUNCH_CHECK();
  return YH_OK;
}

}  // namespace

int yh_rd_rk_supported(const YhK &k) {
  if (k.timeIntOrder != 1 && k.timeIntOrder != 2 && k.timeIntOrder != 4) return 0;
  if (!k.neumannBC || k.anisotropy) return 0;
  if ((k.nx & 1) || k.nx < 16) return 0;
  return 1;
}

int yh_launch_rd_rk(const YhK &k, const double *u_in, const double *v_in, double *u_out,
                    double *v_out, double *vtu, double *vtv, const uint8_t *solid, cudaStream_t st) {
  if (!yh_rd_rk_supported(k)) return YH_ERR_UNSUPPORTED;
  if (k.row1 <= k.row0) return YH_OK;
  RkArgs a{u_in, v_in, u_out, v_out, vtu, vtv, solid, 0};
  const char *force_w = getenv("YH_RK_W");
  const long long cells = (long long)k.nx * (k.row1 - k.row0);
  int W = cells >= (1ll << 23) ? 192 : (cells >= (1ll << 21) ? 128 : 64);
  if (force_w) W = atoi(force_w);
  const bool so = k.solidSwitch != 0;
  const bool lap4 = k.lap4 != 0 && !so;   // the mask branch has no 4th-order terms (:184)
#define YH_RK_W(KK, WW)                                                                       \
  {                                                                                           \
    if (so) return launch<KK, WW, false, true>(k, a, st);                                     \
    return lap4 ? launch<KK, WW, true, false>(k, a, st) : launch<KK, WW, false, false>(k, a, st); \
  }
#define YH_RK_DISPATCH(KK)      \
  if (W == 64) YH_RK_W(KK, 64)  \
  if (W == 128) YH_RK_W(KK, 128) \
  YH_RK_W(KK, 192)
  if (k.timeIntOrder == 4) { YH_RK_DISPATCH(4) }
  if (k.timeIntOrder == 2) { YH_RK_DISPATCH(2) }
  YH_RK_DISPATCH(1)
#undef YH_RK_DISPATCH
#undef YH_RK_W
}
