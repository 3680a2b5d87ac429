// rd_rkq.cu -- second generation of the fused RK4 (+ 4th-order Laplacian) streaming kernel of rd_rk.cu:
// the reference's DEFAULT mode (saveFiles.cu:124-132: timeIntOrder = 4, lap4, gateDiff, no-flux,
// no masks; reactionDiffusion.cu:71-93, 115-141, 186-247, 498-561) with FOUR columns per thread.
//
// Same pipeline as rd_rk.cu -- a CTA owns a strip of W = 128 columns and streams down the rows, level 0
// ("P") canonicalises a raw row and evaluates the currents J = (I_sum, I_v), level s+1 (one warp per RK
// stage s) turns three rows of stage-s state (U, V, Ju, Jv) into du_s, accumulates rhs += w_s du_s and
// emits the stage-(s+1) row with its currents.  P has no warp of its own: the warp of the LAST stage
// does it (that stage forms no next state and no currents, so stage 3 + P costs what the other stages
// cost), which keeps a CTA at four warps -- two CTAs per SM are then two warps per scheduler and the
// kernel may use up to 255 registers (it needs ~220: three rows of four arrays in registers).
// Same lessons as rd_quad.cu, whose header
// explains them: the kernel is bound by ISSUE slots (an FP64 instruction holds the dispatch port two
// cycles, nothing issues in its shadow) and by the shared-memory data pipe (ncu of rd_rk_stream at
// 8192^2: LSU wavefronts 77 % of peak, profiles/r2a_rk4lap4_8192_*), so
//   * a thread owns a QUAD of cells: E/W neighbours inside the quad are registers, the outer ones cost
//     two 8-byte loads per array instead of four, control flow is paid once per four cells;
//   * 3-row rings + a loop unrolled by three: compile-time ring offsets on 32-bit shared addresses;
//   * rows in the 128-byte XOR swizzle: 32-byte quads at stride 32 B are bank-conflict free;
//   * x mirrors are addresses, y mirrors are loads (never register copies);
//   * no loader warp: P fetches its own quads with cp.async, PF rows ahead.
//
//   level l handles row m in iteration m - c0 + 2l (c0 = first raw row of the chunk), reading row m+1 of
//   level l-1, written one iteration earlier.
//
// ARITH: 0 = exact (the reference's expressions, operation for operation, no contraction: bit-identical
// to the plain-C oracle and to rd_rk.cu); 1 = fast (the same update with the stencil coefficients
// combined on the host and FMA chains: ~40 instead of ~79 FP64 instructions per cell and stage; differs
// from exact by rounding only -- pinned by tests/test_gpu_arith.py).
#include <stdlib.h>

#include <type_traits>

#include "rd_euler_cell.cuh"

namespace {

using namespace yh_euler;

struct RkqArgs {
  const double *u_in, *v_in;
  double *u_out, *v_out, *vtu, *vtv;
  int RY;
  // exact flavour: loop-invariant products of the 4th-order terms, formed on the host with the kernel's
  // own operations (rd_rk.cu)
  double q4, m2q, mrs2q4, rsq;
  // fast flavour:  d = cC*C + cH*(W+E) + cV*(N+S) + cQ*(SW+SE+NW+NE) + jC*Jc - jX*(JW+JE) - jY*(JN+JS)
  double uC, uH, uV, uQ, vC, vH, vV, vQ, jC, jX, jY, neg_eps;
};

struct Q4 { double2 a, b; };
struct KRow { Q4 u, v, ju, jv; double uW, uE, vW, vE; };

template <int IMM>
__device__ __forceinline__ double2 lds128(unsigned a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(a), "n"(IMM));
  return v;
}
template <int IMM>
__device__ __forceinline__ double lds64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(IMM));
  return v;
}
template <int IMM>
__device__ __forceinline__ void sts128(unsigned a, const double2 &v) {
  asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(a), "n"(IMM), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ int swz(int ch) { return ((ch >> 3) << 7) | ((((ch & 7) ^ (ch >> 3)) & 7) << 4); }
__device__ __forceinline__ double flip(double t) {   // -(t) on the ALU pipe: bit-identical for every non-NaN t
  return __hiloint2double(__double2hiint(t) ^ (int)0x80000000, __double2loint(t));
}

// Ionic currents (reactionDiffusion.cu:131-141), stimulus off.  DEF: the reference's default constants
// mu = delta = 1, gamma = theta = 0 (1.0*x == x, x - 0.0 == x: dropped without changing a bit).
template <bool DEF, int ARITH>
__device__ __forceinline__ void currents(const YhK &k, const RkqArgs &a, double u, double v, double &ju, double &jv) {
  const double mu_u = DEF ? u : k.mu * u;
  const double ug = DEF ? u : k.delta * (u - k.gamma);
  if (ARITH == 0) {
    ju = flip(mu_u * (1.0 - u) * (u - k.alpha) - u * v);
    const double yv = ug * (k.beta - u) - v;
    jv = -(k.eps * (DEF ? yv : yv - k.theta));
  } else {
    ju = flip(fma(mu_u * (1.0 - u), u - k.alpha, -(u * v)));
    const double yv = fma(ug, k.beta - u, -v);
    jv = a.neg_eps * (DEF ? yv : yv - k.theta);
  }
}

// du of one cell from its 3 x 3 neighbourhood of one field (F = 0: u with Ju, F = 1: v with Jv).
template <int F, bool LAP4, int ARITH>
__device__ __forceinline__ double stage_cell(const YhK &k, const RkqArgs &a, double C, double W, double E, double N,
                                             double S, double SW, double SE, double NW, double NE, double Jc, double JW,
                                             double JE, double JN, double JS) {
  if (ARITH == 0) {   // reactionDiffusion.cu:188-197, 221-239, 498-499, expression for expression
    double d;
    // 2.0*C is exact, so fma(-2.0, C, w) == w - 2.0*C bit for bit
    if (F == 0) d = ((fma(-2.0, C, W) + E) * k.rx + (fma(-2.0, C, N) + S) * k.ry);
    else d = ((fma(-2.0, C, W) + E) * k.rx * k.rscale + (fma(-2.0, C, N) + S) * k.ry * k.rscale);
    if (LAP4) {
      d += (F == 0 ? a.m2q : a.mrs2q4) * (+(W - C + E) + (N - C + S));
      d += (F == 0 ? a.q4 : a.rsq) * (SW + SE + NW + NE);
      d -= ((fma(-2.0, Jc, JW) + JE) * k.fx4 + (fma(-2.0, Jc, JN) + JS) * k.fy4);
    }
    d -= k.dt * Jc;
    return d;
  } else {
    double d = (F == 0 ? a.uC : a.vC) * C;
    d = fma(F == 0 ? a.uH : a.vH, W + E, d);
    d = fma(F == 0 ? a.uV : a.vV, N + S, d);
    if (LAP4) {
      d = fma(F == 0 ? a.uQ : a.vQ, (SW + NW) + (SE + NE), d);
      d = fma(-a.jX, JW + JE, d);
      d = fma(-a.jY, JN + JS, d);
    }
    return fma(a.jC, Jc, d);
  }
}

// Runge-Kutta tables (reactionDiffusion.cu:71-86), stage-indexed: a_{s+1} and w_s of RK4
__constant__ double RKQ_A_NEXT[4] = {0.5, 0.5, 1.0, 0.0};
__constant__ double RKQ_W[4] = {0.166666666666667, 0.333333333333333, 0.333333333333333, 0.166666666666667};

template <bool LAP4, bool DEF, int ARITH>
__global__ void __launch_bounds__(128, 2)
rd_rk_quad(const __grid_constant__ YhK k, const __grid_constant__ RkqArgs a) {
  constexpr int K = 4, W = 128, H = 4;
  constexpr int BX = W - 2 * H;
  constexpr int NL = W / 4;              // 32 threads = one warp per level
  constexpr int FROW = W * 8;            // bytes of one array row
  constexpr int ROW0 = 2 * FROW;         // raw ring row: u0 v0
  constexpr int ROWA = 6 * FROW;         // stage ring row: U V Ju Jv ru rv
  constexpr int NR0 = 16, PF = 5, NRA = 3;
  extern __shared__ __align__(1024) unsigned char smraw[];

  const int tid = threadIdx.x;
  const int st = __shfl_sync(0xffffffffu, tid / NL, 0);    // RK stage of this warp (warp-uniform); warp K-1 is also P
  const int lev = st + 1;                                  // pipeline level of the stage (P = level 0)
  const bool is_p = (st == K - 1);
  const int t = tid % NL;
  const int nx = k.nx;
  const int x0 = blockIdx.x * BX, wx0 = x0 - H;
  const int y0 = k.row0 + blockIdx.y * a.RY;
  const int RYe = min(a.RY, k.row1 - y0);
  const int c0 = y0 - K;
  const int dom_lo = -k.jg0, dom_hi = k.nyg - k.jg0;
  const int n_it = ((RYe + 3 * K + 2) / 3) * 3;

  const int c = 4 * t;
  const int gx = wx0 + c;
  const bool okA = (gx >= 0) && (gx < nx), okB = (gx + 2 >= 0) && (gx + 2 < nx);   // nx % 4 == 0: okA == okB
  const bool col_ok = okA || okB;
  const bool out_col = col_ok && (c >= H) && (c + 4 <= W - H);

  // Static shared-memory offsets of this thread inside an array row.  The outer neighbours of a quad (W of
  // its first cell, E of its last) come from the adjacent lanes by warp shuffle -- the 8-byte loads at
  // stride 32 B they replace were 2-way bank conflicts and a quarter of the kernel's shared-memory
  // wavefronts (ncu: LSU data pipe 83 % busy in the fast flavour).  The no-flux mirror in x stays an
  // ADDRESS: the dead lane left of x = 0 loads the pair (x0, x1) as its second pair, so the lane at x = 0
  // receives x1 as its W value; the dead lane right of x = nx-1 loads (x_{nx-2}, x_{nx-1}) as its first.
  int offA = swz(2 * t), offB = swz(2 * t + 1);   // own quad: every store, and the loads of live lanes
  int ldA = offA, ldB = offB;                       // stage-state loads (aliased for the two dead edge lanes)
  if (gx == -4) ldB = swz(2 * t + 2);
  if (gx == nx && t > 0) ldA = swz(2 * t - 1);
  asm volatile("" : "+r"(offA), "+r"(offB), "+r"(ldA), "+r"(ldB));

  const unsigned sm0 = (unsigned)__cvta_generic_to_shared(smraw);        // raw ring, NR0 rows
  const unsigned smA = sm0 + NR0 * ROW0;                                 // stage rings A_0 .. A_{K-1}
  const unsigned src_ring = smA + st * NRA * ROWA;                       // A_st: input of stage st
  const unsigned dst_ring = smA + (st + 1) * NRA * ROWA;                 // A_{st+1}: output (unused by the last stage)
  const unsigned aA = src_ring + ldA, aB = src_ring + ldB;
  const unsigned dA = dst_ring + offA, dB = dst_ring + offB;
  const unsigned pA = smA + offA, pB = smA + offB;                       // P writes A_0
  const unsigned rA = sm0 + offA, rB = sm0 + offB;

  // rows: P loads / evaluates [c0, y0+RYe+K), stage s produces [y0-(K-1-s), y0+RYe+(K-1-s))
  // (warp-uniform: lanes outside the domain run along -- the shuffles need every lane, and the dead lanes
  // next to x = 0 / x = nx-1 carry the mirror values; only their global stores are masked)
  const int lo_l = max(dom_lo, y0 - (K - 1 - st));
  const int hi_l = min(dom_hi, y0 + RYe + (K - 1 - st));
  const int lo_p = max(dom_lo, c0);
  const int hi_p = min(dom_hi, y0 + RYe + K);
  const int m0 = c0 - 2 * lev;           // row the stage handles in iteration 0
  const double a_next = RKQ_A_NEXT[st], w_k = RKQ_W[st];

  // ---- raw rows: P fetches its own quads, PF rows ahead -------------------------------------------
  const double *pu = a.u_in + gx + (ptrdiff_t)c0 * nx;
  const double *pv = a.v_in + gx + (ptrdiff_t)c0 * nx;
  int szA = okA ? 16 : 0, szB = okB ? 16 : 0;
  asm volatile("" : "+r"(szA), "+r"(szB));
  auto issue_row = [&](int q) {          // called with q = c0, c0+1, ... in order
    if (q >= lo_p && q < hi_p) {
      const unsigned dst = sm0 + (unsigned)((q - c0) & (NR0 - 1)) * ROW0;
      cp_async16s(dst + offA, pu, szA);
      cp_async16s(dst + offB, pu + 2, szB);
      cp_async16s(dst + FROW + offA, pv, szA);
      cp_async16s(dst + FROW + offB, pv + 2, szB);
    }
    pu += nx; pv += nx;
    cp_async_commit();
  };

  // P: stage-0 state u0 + (0.0*0.0) and its currents for row m = c0 + it -> A_0[it % 3]
  auto p_step = [&](auto Jc, int it) {
    constexpr int J = decltype(Jc)::value;
    const int m = c0 + it;
    if (m >= lo_p && m < hi_p) {
      const unsigned r = (unsigned)(it & (NR0 - 1)) * ROW0;
      Q4 u, v, ju, jv;
      u.a = lds128<0>(rA + r); u.b = lds128<0>(rB + r);
      v.a = lds128<FROW>(rA + r); v.b = lds128<FROW>(rB + r);
      u.a.x += 0.0; u.a.y += 0.0; u.b.x += 0.0; u.b.y += 0.0;
      v.a.x += 0.0; v.a.y += 0.0; v.b.x += 0.0; v.b.y += 0.0;
      currents<DEF, ARITH>(k, a, u.a.x, v.a.x, ju.a.x, jv.a.x);
      currents<DEF, ARITH>(k, a, u.a.y, v.a.y, ju.a.y, jv.a.y);
      currents<DEF, ARITH>(k, a, u.b.x, v.b.x, ju.b.x, jv.b.x);
      currents<DEF, ARITH>(k, a, u.b.y, v.b.y, ju.b.y, jv.b.y);
      constexpr int D = (J % NRA) * ROWA;
      sts128<D>(pA, u.a); sts128<D>(pB, u.b);
      sts128<D + FROW>(pA, v.a); sts128<D + FROW>(pB, v.b);
      sts128<D + 2 * FROW>(pA, ju.a); sts128<D + 2 * FROW>(pB, ju.b);
      sts128<D + 3 * FROW>(pA, jv.a); sts128<D + 3 * FROW>(pB, jv.b);
    }
  };

  // one stage-state row (U, V, Ju, Jv and the outer neighbours of U, V) into registers
  auto ld_row = [&](auto Ic, unsigned bA, unsigned bB, KRow &R) {
    constexpr int I = decltype(Ic)::value;
    R.u.a = lds128<I>(bA); R.u.b = lds128<I>(bB);
    R.v.a = lds128<I + FROW>(bA); R.v.b = lds128<I + FROW>(bB);
    R.uW = __shfl_up_sync(0xffffffffu, R.u.b.y, 1); R.uE = __shfl_down_sync(0xffffffffu, R.u.a.x, 1);
    R.vW = __shfl_up_sync(0xffffffffu, R.v.b.y, 1); R.vE = __shfl_down_sync(0xffffffffu, R.v.a.x, 1);
    if (LAP4) {
      R.ju.a = lds128<I + 2 * FROW>(bA); R.ju.b = lds128<I + 2 * FROW>(bB);
      R.jv.a = lds128<I + 3 * FROW>(bA); R.jv.b = lds128<I + 3 * FROW>(bB);
    }
  };
  auto ld_row_dyn = [&](int sb, KRow &R) { ld_row(std::integral_constant<int, 0>{}, aA + sb, aB + sb, R); };

  // "plain" iterations of a stage warp: the row is computed and neither it nor its N row touches a domain edge in y
  const int pl_lo = max(lo_l, dom_lo + 1), pl_hi = min(hi_l, dom_hi - 1);
  const int it_p0 = pl_lo - m0;
  const unsigned n_plain = (pl_hi > pl_lo) ? (unsigned)(pl_hi - pl_lo) : 0u;
  const int it_first = lo_l - 2 - m0;
  const unsigned n_act = (unsigned)(hi_l - lo_l + 2);

  double *gu = a.u_out + gx + (ptrdiff_t)lo_l * nx;   // running output pointers of the last stage
  double *gv = a.v_out + gx + (ptrdiff_t)lo_l * nx;
  double *gtu = a.vtu ? a.vtu + gx + (ptrdiff_t)lo_l * nx : nullptr;
  double *gtv = a.vtv ? a.vtv + gx + (ptrdiff_t)lo_l * nx : nullptr;

  // stage st: row m = m0 + it from rows S (m-1), C (m), N (m+1) of A_st
  auto s_step = [&](auto Jc_, int it, KRow &S, KRow &C, KRow &N) {
    constexpr int J = decltype(Jc_)::value;
    constexpr int SN = ((J + 2) % NRA) * ROWA;   // row m+1: written one iteration ago
    constexpr int SC = ((J + 1) % NRA) * ROWA;   // row m:   two iterations ago
    constexpr int SM = (J % NRA) * ROWA;         // row m-1 (its slot is due for row m+2)
    const bool plain = (unsigned)(it - it_p0) < n_plain;
    bool comp = plain;
    if (plain) {
      ld_row(std::integral_constant<int, SN>{}, aA, aB, N);
    } else if ((unsigned)(it - it_first) < n_act) {   // register fill, first / last row of the domain
      const int m = m0 + it;
      if (m + 1 < dom_hi) {
        if (m + 1 >= dom_lo) ld_row_dyn(SN, N);
      } else {
        ld_row_dyn(SM, N);                      // last row of the domain: N := row m-1
      }
      if (m >= lo_l) {
        comp = true;
        if (m == dom_lo) ld_row_dyn(SN, S);     // first row of the domain: S := row m+1
      }
    }
    if (!comp) return;
    // outer neighbours of the currents of row m, its running rhs, and the raw row
    double juW = 0.0, juE = 0.0, jvW = 0.0, jvE = 0.0;
    if (LAP4) {
      juW = __shfl_up_sync(0xffffffffu, C.ju.b.y, 1); juE = __shfl_down_sync(0xffffffffu, C.ju.a.x, 1);
      jvW = __shfl_up_sync(0xffffffffu, C.jv.b.y, 1); jvE = __shfl_down_sync(0xffffffffu, C.jv.a.x, 1);
    } else {   // without the 4th-order terms only the currents of the row itself are needed
      C.ju.a = lds128<SC + 2 * FROW>(aA); C.ju.b = lds128<SC + 2 * FROW>(aB);
      C.jv.a = lds128<SC + 3 * FROW>(aA); C.jv.b = lds128<SC + 3 * FROW>(aB);
    }
    Q4 ru, rv;
    if (st == 0) {                              // rhs starts as 0.0 (reactionDiffusion.cu:502-503: rhs = 0 + w du)
      ru.a = ru.b = rv.a = rv.b = make_double2(0.0, 0.0);
    } else {
      ru.a = lds128<SC + 4 * FROW>(aA); ru.b = lds128<SC + 4 * FROW>(aB);
      rv.a = lds128<SC + 5 * FROW>(aA); rv.b = lds128<SC + 5 * FROW>(aB);
    }
    const unsigned r0 = (unsigned)((it - 2 * lev) & (NR0 - 1)) * ROW0;
    Q4 u0, v0;
    u0.a = lds128<0>(rA + r0); u0.b = lds128<0>(rB + r0);
    v0.a = lds128<FROW>(rA + r0); v0.b = lds128<FROW>(rB + r0);

    double du[4], dv[4];
    du[0] = stage_cell<0, LAP4, ARITH>(k, a, C.u.a.x, C.uW, C.u.a.y, N.u.a.x, S.u.a.x, S.uW, S.u.a.y, N.uW, N.u.a.y,
                                       C.ju.a.x, juW, C.ju.a.y, N.ju.a.x, S.ju.a.x);
    du[1] = stage_cell<0, LAP4, ARITH>(k, a, C.u.a.y, C.u.a.x, C.u.b.x, N.u.a.y, S.u.a.y, S.u.a.x, S.u.b.x, N.u.a.x, N.u.b.x,
                                       C.ju.a.y, C.ju.a.x, C.ju.b.x, N.ju.a.y, S.ju.a.y);
    du[2] = stage_cell<0, LAP4, ARITH>(k, a, C.u.b.x, C.u.a.y, C.u.b.y, N.u.b.x, S.u.b.x, S.u.a.y, S.u.b.y, N.u.a.y, N.u.b.y,
                                       C.ju.b.x, C.ju.a.y, C.ju.b.y, N.ju.b.x, S.ju.b.x);
    du[3] = stage_cell<0, LAP4, ARITH>(k, a, C.u.b.y, C.u.b.x, C.uE, N.u.b.y, S.u.b.y, S.u.b.x, S.uE, N.u.b.x, N.uE,
                                       C.ju.b.y, C.ju.b.x, juE, N.ju.b.y, S.ju.b.y);
    dv[0] = stage_cell<1, LAP4, ARITH>(k, a, C.v.a.x, C.vW, C.v.a.y, N.v.a.x, S.v.a.x, S.vW, S.v.a.y, N.vW, N.v.a.y,
                                       C.jv.a.x, jvW, C.jv.a.y, N.jv.a.x, S.jv.a.x);
    dv[1] = stage_cell<1, LAP4, ARITH>(k, a, C.v.a.y, C.v.a.x, C.v.b.x, N.v.a.y, S.v.a.y, S.v.a.x, S.v.b.x, N.v.a.x, N.v.b.x,
                                       C.jv.a.y, C.jv.a.x, C.jv.b.x, N.jv.a.y, S.jv.a.y);
    dv[2] = stage_cell<1, LAP4, ARITH>(k, a, C.v.b.x, C.v.a.y, C.v.b.y, N.v.b.x, S.v.b.x, S.v.a.y, S.v.b.y, N.v.a.y, N.v.b.y,
                                       C.jv.b.x, C.jv.a.y, C.jv.b.y, N.jv.b.x, S.jv.b.x);
    dv[3] = stage_cell<1, LAP4, ARITH>(k, a, C.v.b.y, C.v.b.x, C.vE, N.v.b.y, S.v.b.y, S.v.b.x, S.vE, N.v.b.x, N.vE,
                                       C.jv.b.y, C.jv.b.x, jvE, N.jv.b.y, S.jv.b.y);
    // running rhs (:502-503)
    if (ARITH == 0) {
      ru.a.x += (w_k * du[0]); ru.a.y += (w_k * du[1]); ru.b.x += (w_k * du[2]); ru.b.y += (w_k * du[3]);
      rv.a.x += (w_k * dv[0]); rv.a.y += (w_k * dv[1]); rv.b.x += (w_k * dv[2]); rv.b.y += (w_k * dv[3]);
    } else {
      ru.a.x = fma(w_k, du[0], ru.a.x); ru.a.y = fma(w_k, du[1], ru.a.y); ru.b.x = fma(w_k, du[2], ru.b.x); ru.b.y = fma(w_k, du[3], ru.b.y);
      rv.a.x = fma(w_k, dv[0], rv.a.x); rv.a.y = fma(w_k, dv[1], rv.a.y); rv.b.x = fma(w_k, dv[2], rv.b.x); rv.b.y = fma(w_k, dv[3], rv.b.y);
    }
    if (st < K - 1) {
      // stage st+1 state (:117-118), its currents, and the rhs travel to the next warp
      Q4 un, vn, ju, jv;
      if (ARITH == 0) {
        un.a.x = u0.a.x + (a_next * du[0]); un.a.y = u0.a.y + (a_next * du[1]);
        un.b.x = u0.b.x + (a_next * du[2]); un.b.y = u0.b.y + (a_next * du[3]);
        vn.a.x = v0.a.x + (a_next * dv[0]); vn.a.y = v0.a.y + (a_next * dv[1]);
        vn.b.x = v0.b.x + (a_next * dv[2]); vn.b.y = v0.b.y + (a_next * dv[3]);
      } else {
        un.a.x = fma(a_next, du[0], u0.a.x); un.a.y = fma(a_next, du[1], u0.a.y);
        un.b.x = fma(a_next, du[2], u0.b.x); un.b.y = fma(a_next, du[3], u0.b.y);
        vn.a.x = fma(a_next, dv[0], v0.a.x); vn.a.y = fma(a_next, dv[1], v0.a.y);
        vn.b.x = fma(a_next, dv[2], v0.b.x); vn.b.y = fma(a_next, dv[3], v0.b.y);
      }
      currents<DEF, ARITH>(k, a, un.a.x, vn.a.x, ju.a.x, jv.a.x);
      currents<DEF, ARITH>(k, a, un.a.y, vn.a.y, ju.a.y, jv.a.y);
      currents<DEF, ARITH>(k, a, un.b.x, vn.b.x, ju.b.x, jv.b.x);
      currents<DEF, ARITH>(k, a, un.b.y, vn.b.y, ju.b.y, jv.b.y);
      constexpr int D = (J % NRA) * ROWA;
      sts128<D>(dA, un.a); sts128<D>(dB, un.b);
      sts128<D + FROW>(dA, vn.a); sts128<D + FROW>(dB, vn.b);
      sts128<D + 2 * FROW>(dA, ju.a); sts128<D + 2 * FROW>(dB, ju.b);
      sts128<D + 3 * FROW>(dA, jv.a); sts128<D + 3 * FROW>(dB, jv.b);
      sts128<D + 4 * FROW>(dA, ru.a); sts128<D + 4 * FROW>(dB, ru.b);
      sts128<D + 5 * FROW>(dA, rv.a); sts128<D + 5 * FROW>(dB, rv.b);
    } else {
      if (out_col) {
        Q4 uo, vo;   // :512-513
        if (ARITH == 0) {
          uo.a.x = u0.a.x + k.tc * ru.a.x; uo.a.y = u0.a.y + k.tc * ru.a.y; uo.b.x = u0.b.x + k.tc * ru.b.x; uo.b.y = u0.b.y + k.tc * ru.b.y;
          vo.a.x = v0.a.x + k.tc * rv.a.x; vo.a.y = v0.a.y + k.tc * rv.a.y; vo.b.x = v0.b.x + k.tc * rv.b.x; vo.b.y = v0.b.y + k.tc * rv.b.y;
        } else {
          uo.a.x = fma(k.tc, ru.a.x, u0.a.x); uo.a.y = fma(k.tc, ru.a.y, u0.a.y); uo.b.x = fma(k.tc, ru.b.x, u0.b.x); uo.b.y = fma(k.tc, ru.b.y, u0.b.y);
          vo.a.x = fma(k.tc, rv.a.x, v0.a.x); vo.a.y = fma(k.tc, rv.a.y, v0.a.y); vo.b.x = fma(k.tc, rv.b.x, v0.b.x); vo.b.y = fma(k.tc, rv.b.y, v0.b.y);
        }
        *reinterpret_cast<double2 *>(gu) = uo.a; *reinterpret_cast<double2 *>(gu + 2) = uo.b;
        *reinterpret_cast<double2 *>(gv) = vo.a; *reinterpret_cast<double2 *>(gv + 2) = vo.b;
        if (gtu) {   // velTan = rhs / dt (:529-530, 551-552)
          double2 ta, tb;
          ta.x = ru.a.x / k.dt; ta.y = ru.a.y / k.dt; tb.x = ru.b.x / k.dt; tb.y = ru.b.y / k.dt;
          *reinterpret_cast<double2 *>(gtu) = ta; *reinterpret_cast<double2 *>(gtu + 2) = tb;
          ta.x = rv.a.x / k.dt; ta.y = rv.a.y / k.dt; tb.x = rv.b.x / k.dt; tb.y = rv.b.y / k.dt;
          *reinterpret_cast<double2 *>(gtv) = ta; *reinterpret_cast<double2 *>(gtv + 2) = tb;
        }
      }
      gu += nx; gv += nx;
      if (gtu) { gtu += nx; gtv += nx; }
    }
  };

  KRow RA, RB, RC;
  RA.u.a = RA.u.b = RA.v.a = RA.v.b = RA.ju.a = RA.ju.b = RA.jv.a = RA.jv.b = make_double2(0, 0);
  RA.uW = RA.uE = RA.vW = RA.vE = 0.0;
  RB = RA; RC = RA;

  if (is_p) {
#pragma unroll
    for (int q = 0; q < PF; q++) issue_row(c0 + q);
  }
  auto sub = [&](auto Jc, int it0, KRow &S, KRow &C, KRow &N) {
    constexpr int J = decltype(Jc)::value;
    const int it = it0 + J;
    if (is_p) {
      issue_row(c0 + PF + it);
      cp_async_wait<PF>();               // this thread's pieces of rows <= c0 + it have landed
    }
    __syncthreads();                     // rows written in the last iteration (rings and raw ring) are visible
    if (is_p) p_step(Jc, it);
    s_step(Jc, it, S, C, N);
  };
  for (int it = 0; it < n_it; it += 3) {
    sub(std::integral_constant<int, 0>{}, it, RA, RB, RC);
    sub(std::integral_constant<int, 1>{}, it, RB, RC, RA);
    sub(std::integral_constant<int, 2>{}, it, RC, RA, RB);
  }
  if (is_p) cp_async_wait<0>();
}

template <bool LAP4, bool DEF, int ARITH>
int launch_q(const YhK &k, RkqArgs a, cudaStream_t st) {
  constexpr int W = 128, BX = W - 8, FROW = W * 8;
  constexpr int NT = 128;
  const size_t smem = (size_t)16 * 2 * FROW + (size_t)4 * 3 * 6 * FROW;
  static bool attr_set[64] = {false};
  static int slots[64] = {0};
  int dev = 0;
  YH_CUDA(cudaGetDevice(&dev));
  auto kfn = rd_rk_quad<LAP4, DEF, ARITH>;
  if (!attr_set[dev & 63]) {
    YH_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1, sms = 148;
    YH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, NT, smem));
    YH_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    slots[dev & 63] = (per_sm > 0 ? per_sm : 1) * sms;
    attr_set[dev & 63] = true;
  }
  const int rows = k.row1 - k.row0;
  const int strips = (k.nx + BX - 1) / BX;
  const char *force_ry = getenv("YH_RK_RY");
  a.RY = force_ry ? atoi(force_ry) : pick_ry_waves(rows, strips, 1, 4, slots[dev & 63]);
  dim3 grd(strips, (rows + a.RY - 1) / a.RY);
  YH_LAUNCH(kfn, grd, NT, smem, st, k, a);
  return YH_OK;
}

}  // namespace

// The reference's default switches, no masks, a sheet wide enough for 128-column strips.
int yh_rd_rkq_supported(const YhK &k) {
  if (k.timeIntOrder != 4 || !k.neumannBC || k.anisotropy || k.solidSwitch) return 0;
  if (!k.gateDiff || k.stim) return 0;
  if ((k.nx & 3) || k.nx < 128) return 0;
  return 1;
}

int yh_launch_rd_rkq(const YhK &k, int arith, const double *u_in, const double *v_in, double *u_out, double *v_out,
                     double *vtu, double *vtv, cudaStream_t st) {
  if (!yh_rd_rkq_supported(k)) return YH_ERR_UNSUPPORTED;
  if (k.row1 <= k.row0) return YH_OK;
  RkqArgs a;
  a.u_in = u_in; a.v_in = v_in; a.u_out = u_out; a.v_out = v_out; a.vtu = vtu; a.vtv = vtv; a.RY = 0;
  const bool lap4 = k.lap4 != 0;
  {
    volatile double q4 = k.qx4 + k.qy4;          // volatile: every product is rounded to double here
    volatile double m2q = -2.0 * q4;
    volatile double mrs2 = -k.rscale * 2.0;
    volatile double mrs2q4 = mrs2 * q4;
    volatile double rsq = k.rscale * q4;
    a.q4 = q4; a.m2q = m2q; a.mrs2q4 = mrs2q4; a.rsq = rsq;
  }
  {   // fast flavour: the same linear combination with its coefficients collected
    const double q = lap4 ? k.qx4 + k.qy4 : 0.0;
    a.uH = k.rx - 2.0 * q; a.uV = k.ry - 2.0 * q; a.uC = -2.0 * (k.rx + k.ry) + 4.0 * q; a.uQ = q;
    a.vH = k.rscale * a.uH; a.vV = k.rscale * a.uV; a.vC = k.rscale * a.uC; a.vQ = k.rscale * a.uQ;
    a.jX = lap4 ? k.fx4 : 0.0; a.jY = lap4 ? k.fy4 : 0.0;
    a.jC = (lap4 ? 2.0 * (k.fx4 + k.fy4) : 0.0) - k.dt;
    a.neg_eps = -k.eps;
  }
  const bool def = (k.mu == 1.0) && (k.delta == 1.0) && (k.gamma == 0.0) && (k.theta == 0.0);
#define YH_RKQ(L, D)                                                     \
  return arith ? launch_q<L, D, 1>(k, a, st) : launch_q<L, D, 0>(k, a, st);
  if (lap4) { if (def) { YH_RKQ(true, true) } else { YH_RKQ(true, false) } }
  if (def) { YH_RKQ(false, true) }
  YH_RKQ(false, false)
#undef YH_RKQ
}
