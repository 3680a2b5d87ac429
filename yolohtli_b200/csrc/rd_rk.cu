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
  // loop-invariant products of the 4th-order terms, formed on the host with the kernel's own
  // operations (IEEE, no contraction): in registers they were rematerialised every row (ncu)
  double q4;              // qx4 + qy4
  double m2q;             // -2.0*( qx4+qy4 )                    (:221)
  double mrs2q4;          // (-rscale*2.0)*( qx4+qy4 )           (:235)
  double rsq;             // rscale*( qx4+qy4 )                  (:239)
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
__device__ __forceinline__ void cp_async16(unsigned smem, const void *gmem, bool valid) {
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem), "l"(gmem), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// Ionic currents (reactionDiffusion.cu:131-141).  DEF = the reference's default constants
// mu = delta = 1, gamma = theta = 0: 1.0*x == x and x - 0.0 == x exactly, so those operations
// are dropped without changing a bit.
// NOSTIM: the caller knows the stimulus is off; the negation is then the last operation before the
// value goes to shared memory, and flipping the sign bit (ALU pipe) gives the bits of -(x) for every
// non-NaN x without an instruction on the FP64 pipe (ptxas emits DADD -RZ, -x otherwise).
template <bool DEF, bool NOSTIM>
__device__ __forceinline__ double rk_Isum(const YhK &k, double u, double v, bool scs) {
  const double mu_u = DEF ? u : k.mu * u;
  const double t = mu_u * (1.0 - u) * (u - k.alpha) - u * v;
  if (NOSTIM) return __hiloint2double(__double2hiint(t) ^ (int)0x80000000, __double2loint(t));
  const double I = -t;
  return scs ? I - 24.7 : I;   // x - 0.0 == x
}
template <bool DEF>
__device__ __forceinline__ double rk_Iv(const YhK &k, double u, double v) {
  const double ug = DEF ? u : k.delta * (u - k.gamma);
  const double yv = ug * (k.beta - u) - v;
  return -(k.eps * (DEF ? yv : yv - k.theta));
}

// Runge-Kutta tables (reactionDiffusion.cu:71-86) as the reference's literals: a_{k+1} (weight of
// du_k in the next stage state) and w_k, stage-indexed, for RK4 | RK2 | Euler.  In constant memory:
// an indexed local array lands in local memory (ncu: 5.9 M local loads per launch), and selecting
// the literals arithmetically from the stage index cost ~10 FSEL/ISETP per row (ncu source page).
__constant__ double RK_A_NEXT[8] = {0.5, 0.5, 1.0, 0.0, /*RK2*/ 0.5, 0.0, /*Euler*/ 0.0, 0.0};
__constant__ double RK_W[8] = {0.166666666666667, 0.333333333333333, 0.333333333333333, 0.166666666666667,
                               /*RK2*/ 0.0, 1.0, /*Euler*/ 1.0, 0.0};

// K stages, strip of W columns.  Arrays of one ring row: U V Ju Jv ru rv (6 x PITCH doubles).
// FAST: gateDiff on and the live stimulus off, known at compile time (the reference's default mode
// is instantiated this way): the u and v halves of a stage and the next stage's currents then form
// ONE basic block -- ptxas does not schedule across the (uniform) branches, and ncu showed the v
// half and the current evaluation as serial dependency chains.  FAST = false reads both switches
// from the parameters.
template <int K, int W, bool LAP4, bool SOLID, bool DEF, bool FAST>
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
  const int n_it = ((RYe + 3 * K + 1 + 2) / 3) * 3;   // multiple of the unroll factor

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

  const int st = g - 1;                  // stage index of this group (-1: group P)
  const int rk_i = (K == 4 ? 0 : (K == 2 ? 4 : 6)) + (st > 0 ? st : 0);
  const double a_next = RK_A_NEXT[rk_i], w_k = RK_W[rk_i];
  const bool gd = FAST ? true : (k.gateDiff != 0);

  // rows each group handles (empty for out-of-domain columns)
  int lo_g, hi_g;
  if (g == 0) { lo_g = max(dom_lo, y0 - K); hi_g = min(dom_hi, y0 + RYe + K); }
  else { lo_g = max(dom_lo, y0 - (K - 1 - st)); hi_g = min(dom_hi, y0 + RYe + (K - 1 - st)); }
  if (!col_ok) hi_g = lo_g;
  const int m_shift = (g == 0) ? 1 : 1 + 2 * g;   // row m = it + c0 - m_shift

  const double q4 = a.q4, m2q = a.m2q, mrs2q4 = a.mrs2q4, rsq = a.rsq;

  // Source rows m-1, m, m+1 of this group's stage live in registers and rotate by renaming (loop
  // unrolled by three): per iteration only the new N row is read from shared memory -- the
  // kernel was LSU-wavefront bound when it re-read all three rows (ncu: 83 % of the LSU data pipe).
  struct RowRegs { double2 u, v, ju, jv; double uw, ue, vw, ve; };
  const double *Ak = A + (st > 0 ? st : 0) * NRA * ROWA + cc;

  auto load_row = [&](int row, RowRegs &R) {
    const double *p = Ak + ((row - c0) & (NRA - 1)) * ROWA;
    R.u = *reinterpret_cast<const double2 *>(p);
    R.v = *reinterpret_cast<const double2 *>(p + PITCH);
    // W of x=0 / E of x=nx-1: the producer of the row stored the mirror value (x=1 / x=nx-2) in the
    // pad column (store_pads), so no select sits between this load and the stencil
    R.uw = p[-1]; R.ue = p[2]; R.vw = p[PITCH - 1]; R.ve = p[PITCH + 2];
    if (LAP4) {
      R.ju = *reinterpret_cast<const double2 *>(p + 2 * PITCH);
      R.jv = *reinterpret_cast<const double2 *>(p + 3 * PITCH);
    }
  };

  // No-flux mirror in x, once per produced row instead of once per read: the thread that owns x=0
  // also writes its x=1 values into the column of x=-1 (owned by an out-of-domain, idle thread), the
  // one that owns x=nx-1 its x=nx-2 values into the column of x=nx.
  auto store_pads = [&](double *d, const double2 &u, const double2 &v, const double2 &ju, const double2 &jv) {
    if (left_edge) { d[-1] = u.y; d[PITCH - 1] = v.y; d[2 * PITCH - 1] = ju.y; d[3 * PITCH - 1] = jv.y; }
    if (right_edge) { d[2] = u.x; d[PITCH + 2] = v.x; d[2 * PITCH + 2] = ju.x; d[3 * PITCH + 2] = jv.x; }
  };

  auto p_step = [&](int m) {   // ---- P: stage-0 state u0 + (0.0*0.0) and its currents ----
    const int gj = m + k.jg0;
    const double *r0 = R0 + ((m - c0) & (NR0 - 1)) * ROW0 + cc;
    double2 u = *reinterpret_cast<const double2 *>(r0);
    double2 v = *reinterpret_cast<const double2 *>(r0 + PITCH);
    u.x += 0.0; u.y += 0.0; v.x += 0.0; v.y += 0.0;
    double2 ju, jv;
    ju.x = rk_Isum<DEF, FAST>(k, u.x, v.x, FAST ? false : yh_scs(k, gx, gj));
    ju.y = rk_Isum<DEF, FAST>(k, u.y, v.y, FAST ? false : yh_scs(k, gx + 1, gj));
    jv.x = rk_Iv<DEF>(k, u.x, v.x);
    jv.y = rk_Iv<DEF>(k, u.y, v.y);
    double *d = A + ((m - c0) & (NRA - 1)) * ROWA + cc;
    *reinterpret_cast<double2 *>(d) = u;
    *reinterpret_cast<double2 *>(d + PITCH) = v;
    *reinterpret_cast<double2 *>(d + 2 * PITCH) = ju;
    *reinterpret_cast<double2 *>(d + 3 * PITCH) = jv;
    store_pads(d, u, v, ju, jv);
    // the running rhs starts as 0.0 in the ring, so every stage group reads it the same way (no
    // per-row select between "first stage" and the others in the heavy groups; P has cycles to spare)
    *reinterpret_cast<double2 *>(d + 4 * PITCH) = make_double2(0.0, 0.0);
    *reinterpret_cast<double2 *>(d + 5 * PITCH) = make_double2(0.0, 0.0);
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
  };

  // ---- S_st: du of stage st for the pair (c, c+1) of row m from rows S (m-1), C (m), N (m+1) ----
  auto s_step = [&](int m, RowRegs &S, RowRegs &C, RowRegs &N) {
    const int slot = (m - c0) & (NRA - 1);
    const int gj = m + k.jg0;
    if (m == lo_g) {                      // first row of this group: nothing in registers yet
      load_row(m, C);
      load_row((m - 1 < dom_lo) ? m + 1 : m - 1, S);     // no-flux mirror at the first row
    }
    load_row((m + 1 >= dom_hi) ? m - 1 : m + 1, N);      // ... and at the last row
    const double *pc = Ak + slot * ROWA;
    unsigned codes = 0;
    if (SOLID) codes = *reinterpret_cast<const unsigned *>(Cr + ((m - c0) & (NR0 - 1)) * W + c);
    // (an all-tissue fast path per warp -- vote on the codes, plain stencil when every lane is (1,2,1) / (1,2,1) -- was
    // measured on the hole masks of C2 and dropped: 1024^2 49.2 -> 50.9 us, 2048^2 195 -> 207 us per step)
    double du[2], dv[2];
#pragma unroll
    for (int f = 0; f < 2; f++) {   // f = 0: u with Ju, f = 1: v with Jv
      const double2 Cc = f ? C.v : C.u, Ss = f ? S.v : S.u, Nn = f ? N.v : N.u;
      const double Wv = f ? C.vw : C.uw, Ev = f ? C.ve : C.ue;
      double d0, d1;
      if (SOLID) {   // reactionDiffusion.cu:171-180
        // The coefficient triples are exact 0 / 1 / 2 and only four occur per axis for a tissue cell:
        // (1,2,1), (2,2,0), (0,2,2), (0,0,0).  Each equals -- bit for bit, for finite fields -- the
        // plain stencil fma(-2, c, A) + B on substituted neighbours (table in rd_fast.cu), which
        // trades three DMULs per axis for one DADD.  cL = coefficient of the first neighbour.
        const unsigned k0 = codes & 0xFFFFu, k1 = codes >> 16;
        auto axis = [](unsigned cL, unsigned cC, double c, double L, double R) {
          const bool both = cL == 1u;
          const double t = (cL == 2u) ? L : R;
          const double r = fma(-2.0, c, both ? L : t + t) + (both ? R : 0.0);
          return cC ? r : 0.0;
        };
        const double x0 = axis(k0 & 3u, (k0 >> 2) & 3u, Cc.x, Wv, Cc.y), y0 = axis((k0 >> 6) & 3u, (k0 >> 8) & 3u, Cc.x, Nn.x, Ss.x);
        const double x1 = axis(k1 & 3u, (k1 >> 2) & 3u, Cc.y, Cc.x, Ev), y1 = axis((k1 >> 6) & 3u, (k1 >> 8) & 3u, Cc.y, Nn.y, Ss.y);
        if (f == 0) {
          d0 = (x0 * k.rx + y0 * k.ry);
          d1 = (x1 * k.rx + y1 * k.ry);
        } else if (gd) {
          d0 = (x0 * k.rx * k.rscale + y0 * k.ry * k.rscale);
          d1 = (x1 * k.rx * k.rscale + y1 * k.ry * k.rscale);
        } else {
          d0 = 0.0; d1 = 0.0;
        }
      } else if (f == 0) {
        // 2.0*C is exact, so fma(-2.0, C, w) == w - 2.0*C bit for bit
        d0 = ((fma(-2.0, Cc.x, Wv) + Cc.y) * k.rx + (fma(-2.0, Cc.x, Nn.x) + Ss.x) * k.ry);
        d1 = ((fma(-2.0, Cc.y, Cc.x) + Ev) * k.rx + (fma(-2.0, Cc.y, Nn.y) + Ss.y) * k.ry);
      } else if (gd) {
        d0 = ((fma(-2.0, Cc.x, Wv) + Cc.y) * k.rx * k.rscale + (fma(-2.0, Cc.x, Nn.x) + Ss.x) * k.ry * k.rscale);
        d1 = ((fma(-2.0, Cc.y, Cc.x) + Ev) * k.rx * k.rscale + (fma(-2.0, Cc.y, Nn.y) + Ss.y) * k.ry * k.rscale);
      } else {
        d0 = 0.0; d1 = 0.0;
      }
      double2 Jc;
      if (LAP4) Jc = f ? C.jv : C.ju;
      else Jc = *reinterpret_cast<const double2 *>(pc + (2 + f) * PITCH);
      if (LAP4 && (f == 0 || gd)) {
        const double SWv = f ? S.vw : S.uw, SEv = f ? S.ve : S.ue;
        const double NWv = f ? N.vw : N.uw, NEv = f ? N.ve : N.ue;
        const double2 Js = f ? S.jv : S.ju, Jn = f ? N.jv : N.ju;
        const double JW = pc[(2 + f) * PITCH - 1], JE = pc[(2 + f) * PITCH + 2];   // mirrored by store_pads
        if (f == 0) {   // reactionDiffusion.cu:221-229
          d0 += m2q * (+(Wv - Cc.x + Cc.y) + (Nn.x - Cc.x + Ss.x));
          d1 += m2q * (+(Cc.x - Cc.y + Ev) + (Nn.y - Cc.y + Ss.y));
          d0 += q4 * (SWv + Ss.y + NWv + Nn.y);
          d1 += q4 * (Ss.x + SEv + Nn.x + NEv);
        } else {        // :235-239
          d0 += mrs2q4 * (+(Wv - Cc.x + Cc.y) + (Nn.x - Cc.x + Ss.x));
          d1 += mrs2q4 * (+(Cc.x - Cc.y + Ev) + (Nn.y - Cc.y + Ss.y));
          d0 += rsq * (SWv + Ss.y + NWv + Nn.y);
          d1 += rsq * (Ss.x + SEv + Nn.x + NEv);
        }
        d0 -= ((fma(-2.0, Jc.x, JW) + Jc.y) * k.fx4 + (fma(-2.0, Jc.x, Jn.x) + Js.x) * k.fy4);
        d1 -= ((fma(-2.0, Jc.y, Jc.x) + JE) * k.fx4 + (fma(-2.0, Jc.y, Jn.y) + Js.y) * k.fy4);
      }
      d0 -= k.dt * Jc.x;   // :498-499
      d1 -= k.dt * Jc.y;
      if (f == 0) { du[0] = d0; du[1] = d1; }
      else { dv[0] = d0; dv[1] = d1; }
    }
    // running rhs (:502-503)
    double2 ru = *reinterpret_cast<const double2 *>(pc + 4 * PITCH);   // 0.0 from P for the first stage
    double2 rv = *reinterpret_cast<const double2 *>(pc + 5 * PITCH);
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
      ju.x = rk_Isum<DEF, FAST>(k, un.x, vn.x, FAST ? false : yh_scs(k, gx, gj));
      ju.y = rk_Isum<DEF, FAST>(k, un.y, vn.y, FAST ? false : yh_scs(k, gx + 1, gj));
      jv.x = rk_Iv<DEF>(k, un.x, vn.x);
      jv.y = rk_Iv<DEF>(k, un.y, vn.y);
      double *d = A + (st + 1) * NRA * ROWA + slot * ROWA + cc;
      *reinterpret_cast<double2 *>(d) = un;
      *reinterpret_cast<double2 *>(d + PITCH) = vn;
      *reinterpret_cast<double2 *>(d + 2 * PITCH) = ju;
      *reinterpret_cast<double2 *>(d + 3 * PITCH) = jv;
      *reinterpret_cast<double2 *>(d + 4 * PITCH) = ru;
      *reinterpret_cast<double2 *>(d + 5 * PITCH) = rv;
      store_pads(d, un, vn, ju, jv);
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
      if (a.vtu && gd) {   // :551-552
        double2 tu, tv;
        tu.x = sc0 ? ru.x / k.dt : 0.0; tu.y = sc1 ? ru.y / k.dt : 0.0;   // :529-530
        tv.x = sc0 ? rv.x / k.dt : 0.0; tv.y = sc1 ? rv.y / k.dt : 0.0;
        *reinterpret_cast<double2 *>(a.vtu + o) = tu;
        *reinterpret_cast<double2 *>(a.vtv + o) = tv;
      }
    }
  };

  auto sub = [&](int it, RowRegs &S, RowRegs &C, RowRegs &N) {
    const int m = it + c0 - m_shift;
    if (m >= lo_g && m < hi_g) {
      if (g == 0) p_step(m);
      else s_step(m, S, C, N);
    }
    __syncthreads();
  };

  RowRegs RA, RB, RC;
  RA.u = RA.v = RA.ju = RA.jv = make_double2(0, 0);
  RA.uw = RA.ue = RA.vw = RA.ve = 0.0;
  RB = RA; RC = RA;
  for (int it = 0; it < n_it; it += 3) {
    sub(it, RA, RB, RC);
    sub(it + 1, RB, RC, RA);
    sub(it + 2, RC, RA, RB);
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

template <int K, int W, bool LAP4, bool SOLID, bool DEF, bool FAST>
int launch2(const YhK &k, RkArgs a, cudaStream_t st) {
  constexpr int PITCH = W + 4, BX = W - 2 * ((K + 1) & ~1);
  constexpr int NT = (K + 1) * (W / 2) + 32;
  const size_t smem = ((size_t)16 * 2 * PITCH + (size_t)K * 4 * 6 * PITCH) * sizeof(double) +
                      (size_t)16 * W * sizeof(unsigned short);
  static bool attr_set[64] = {false};
  static int slots[64] = {0};
  int dev = 0;
  YH_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    YH_CUDA(cudaFuncSetAttribute(rd_rk_stream<K, W, LAP4, SOLID, DEF, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1, sms = 148;
    YH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rd_rk_stream<K, W, LAP4, SOLID, DEF, FAST>, NT, smem));
    YH_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    slots[dev & 63] = (per_sm > 0 ? per_sm : 1) * sms;
    attr_set[dev & 63] = true;
  }
  const int rows = k.row1 - k.row0;
  const int strips = (k.nx + BX - 1) / BX;
  const char *force_ry = getenv("YH_RK_RY");
  a.RY = force_ry ? atoi(force_ry) : pick_ry_rk(rows, strips, K, slots[dev & 63]);
  dim3 grd(strips, (rows + a.RY - 1) / a.RY);
  auto kfn = rd_rk_stream<K, W, LAP4, SOLID, DEF, FAST>;
  YH_LAUNCH(kfn, grd, NT, smem, st, k, a);
  return YH_OK;
}

template <int K, int W, bool LAP4, bool SOLID>
int launch(const YhK &k, RkArgs a, cudaStream_t st) {
  const bool def = (k.mu == 1.0) && (k.delta == 1.0) && (k.gamma == 0.0) && (k.theta == 0.0);
  if (K == 4 && k.gateDiff && !k.stim) {   // the reference's default switches (with or without masks / lap4)
    constexpr bool F = (K == 4);           // (RK2 / Euler never instantiate FAST)
    return def ? launch2<K, W, LAP4, SOLID, true, F>(k, a, st) : launch2<K, W, LAP4, SOLID, false, F>(k, a, st);
  }
  return def ? launch2<K, W, LAP4, SOLID, true, false>(k, a, st) : launch2<K, W, LAP4, SOLID, false, false>(k, a, st);
}

}  // namespace

int yh_rd_rk_supported(const YhK &k) {
  if (k.timeIntOrder != 1 && k.timeIntOrder != 2 && k.timeIntOrder != 4) return 0;
  if (!k.neumannBC || k.anisotropy) return 0;
  if ((k.nx & 1) || k.nx < 16) return 0;
  return 1;
}

int yh_rd_rkq_supported(const YhK &k);   // rd_rkq.cu: four columns per thread, exact | fast arithmetic
int yh_launch_rd_rkq(const YhK &k, int arith, const double *u_in, const double *v_in, double *u_out, double *v_out,
                     double *vtu, double *vtv, cudaStream_t st);

int yh_launch_rd_rk(const YhK &k, const double *u_in, const double *v_in, double *u_out,
                    double *v_out, double *vtu, double *vtv, const uint8_t *solid, cudaStream_t st) {
  if (!yh_rd_rk_supported(k)) return YH_ERR_UNSUPPORTED;
  if (k.row1 <= k.row0) return YH_OK;
  {   // the reference's default mode on sheets that fill the machine: rd_rkq.cu; YH_RK_KERNEL = quad | pair overrides
    const char *kern = getenv("YH_RK_KERNEL");
    const long long cells_ = (long long)k.nx * (k.row1 - k.row0);
    const bool quad = kern ? (kern[0] == 'q') : (cells_ >= (1ll << 21));
    if (quad && yh_rd_rkq_supported(k))
      return yh_launch_rd_rkq(k, yh_arithmetic(), u_in, v_in, u_out, v_out, (vtu && k.gateDiff) ? vtu : nullptr,
                              (vtu && k.gateDiff) ? vtv : nullptr, st);
  }
  {   // obstacle masks (C2: 1024^2 + holes): the marching tiles, which derive a cell's mask code once per launch and
      // keep it in a register, on the sheets the tile kernels serve.  Measured (B200, RK4 + holes, us per step, march |
      // this kernel): 512^2 14.0 | 18.6, 1024^2 53.5 | 49.2, 4096^2 876 | 624; YH_SOLID_RK = march | stream overrides
    const char *kern = getenv("YH_SOLID_RK");
    const long long cells_ = (long long)k.nx * (k.row1 - k.row0);
    const bool march = kern ? (kern[0] == 'm') : (yh_rd_prefer_tile(cells_) != 0);
    if (march && k.solidSwitch && solid && k.gateDiff && !k.stim && yh_rd_tile_march_solid_supported(k))
      return yh_launch_rd_tile_march_solid(k, u_in, v_in, u_out, v_out, vtu, vtv, solid, st);
  }
  RkArgs a{u_in, v_in, u_out, v_out, vtu, vtv, solid, 0, 0.0, 0.0, 0.0, 0.0};
  {
    volatile double q4 = k.qx4 + k.qy4;          // volatile: every product is rounded to double here
    volatile double m2q = -2.0 * q4;
    volatile double mrs2 = -k.rscale * 2.0;
    volatile double mrs2q4 = mrs2 * q4;
    volatile double rsq = k.rscale * q4;
    a.q4 = q4; a.m2q = m2q; a.mrs2q4 = mrs2q4; a.rsq = rsq;
  }
  const char *force_w = getenv("YH_RK_W");
  const long long cells = (long long)k.nx * (k.row1 - k.row0);
  int W = cells >= (1ll << 23) ? 192 : (cells >= (1ll << 21) ? 128 : 64);
  const bool so = k.solidSwitch != 0;
  // masks: the widest strips pay earlier (measured, RK4 + holes, us per step, W = 64 | 128 | 192: 1024^2 49.2 | 47.6 | 46.7,
  // 2048^2 - | 192 | 168)
  if (so && cells >= (3ll << 18)) W = 192;
  if (force_w) W = atoi(force_w);
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
