// rd_tile.cu -- the monodomain step for SMALL sheets (the reference's own default case is a
// single 512 x 512 sheet, saveFiles.cu:137-138): the streaming kernels of rd_fast.cu / rd_rk.cu
// need hundreds of rows per CTA to amortise their pipeline fill, and a 512^2 sheet cannot give
// 148 SMs that many.  Here a CTA owns a 32 x 32 output tile plus halo entirely in shared memory
// and performs either T Euler steps or the K synchronous Runge-Kutta stages of ONE step with a
// __syncthreads between them; latency per launch is a handful of stage times instead of
// (rows + fill) pipeline iterations.  Same arithmetic, same bits as the streaming kernels.
//
//   rd_tile_rk<K,LAP4>   RK2 / RK4 (+4th-order Laplacian), no-flux, square domain
//                        reactionDiffusion.cu:71-93,115-247,498-561
//   rd_tile_euler<T>     T Euler + 5-point steps per launch (temporal blocking in the tile)
//
// Threads own PAIRS of cells (16-byte shared-memory accesses); the mapping cell -> (thread, slot)
// is fixed for the whole launch so du / rhs of a cell stay in registers across stages.
#include <stdlib.h>

#include "yh_common.cuh"

namespace {

constexpr int TB = 32;          // output tile edge
constexpr int NTHR = 416;        // RK kernel: 13 warps, 2 pairs per thread (832 slots for 800 pairs), 2 CTAs/SM

// NOSTIM: stimulus known to be off; the negation is then the last operation before the value goes to
// shared memory and is done on the sign bit (ALU pipe) instead of the FP64 pipe (see rd_rk.cu).
template <bool DEF, bool NOSTIM>
__device__ __forceinline__ double t_Isum(const YhK &k, double u, double v, bool scs) {
  const double mu_u = DEF ? u : k.mu * u;
  const double t = mu_u * (1.0 - u) * (u - k.alpha) - u * v;
  if (NOSTIM) return __hiloint2double(__double2hiint(t) ^ (int)0x80000000, __double2loint(t));
  const double I = -t;
  return scs ? I - 24.7 : I;   // x - 0.0 == x
}
template <bool DEF>
__device__ __forceinline__ double t_Iv(const YhK &k, double u, double v) {
  const double ug = DEF ? u : k.delta * (u - k.gamma);
  const double yv = ug * (k.beta - u) - v;
  return -(k.eps * (DEF ? yv : yv - k.theta));
}

// A pair (tx, tx+1) is computed at a level whose valid region is the tile minus `ring` outer
// rings when AT LEAST ONE of its cells is inside that region: with odd rings the region's x
// bounds are odd and the boundary cell shares its pair with an invalid one.
__device__ __forceinline__ bool pair_in_ring(int tx, int ty, int ring, int TW) {
  return tx + 1 >= ring && tx < TW - ring && ty >= ring && ty < TW - ring;
}

struct TileArgs {
  const double *u_in, *v_in;
  double *u_out, *v_out, *vtu, *vtv;
  long long sim_stride;
  const int *period;
  int duration, count0;
  // electrode trace fused into the Euler kernel (yh_sim_run with a trace): sample slot0 + s of sheet
  // z = (u, v) at (k.px, k.py) after s of this launch's steps, s = 0 .. T-1; slot0 = *slot
  double *trace;
  const unsigned long long *slot;
  // FAST arithmetic flavour of the RK kernel (yh_set_arithmetic; coefficients as in rd_rkq.cu):
  //   d = cC*C + cH*(W+E) + cV*(N+S) + cQ*(SW+SE+NW+NE) + jC*Jc - jX*(JW+JE) - jY*(JN+JS)
  double uC, uH, uV, uQ, vC, vH, vV, vQ, jC, jX, jY, neg_eps;
};

__device__ __forceinline__ double t_flip(double t) {
  return __hiloint2double(__double2hiint(t) ^ (int)0x80000000, __double2loint(t));
}
// currents of the FAST flavour (stimulus off): FMA-contracted forms of t_Isum / t_Iv
template <bool DEF>
__device__ __forceinline__ void t_currents_fast(const YhK &k, const TileArgs &a, double u, double v, double &ju, double &jv) {
  const double mu_u = DEF ? u : k.mu * u;
  const double ug = DEF ? u : k.delta * (u - k.gamma);
  ju = t_flip(fma(mu_u * (1.0 - u), u - k.alpha, -(u * v)));
  const double yv = fma(ug, k.beta - u, -v);
  jv = a.neg_eps * (DEF ? yv : yv - k.theta);
}
template <int F, bool LAP4>
__device__ __forceinline__ double t_cell_fast(const TileArgs &a, double C, double WE, double NS, double Q, double Jc,
                                              double JWE, double JNS) {
  double d = (F == 0 ? a.uC : a.vC) * C;
  d = fma(F == 0 ? a.uH : a.vH, WE, d);
  d = fma(F == 0 ? a.uV : a.vV, NS, d);
  if (LAP4) {
    d = fma(F == 0 ? a.uQ : a.vQ, Q, d);
    d = fma(-a.jX, JWE, d);
    d = fma(-a.jY, JNS, d);
  }
  return fma(a.jC, Jc, d);
}

// ------------------------------------------------------------------------------------------
// Runge-Kutta tile kernel
// ------------------------------------------------------------------------------------------
// FAST: gateDiff on and the live stimulus off at compile time (the reference's default mode): no
// uniform branches inside a stage, so ptxas schedules the u and v halves as one block (rd_rk.cu).
// ARITH: 0 = exact (the reference's expressions, no contraction), 1 = fast (FAST switches only; see rd_rkq.cu).
template <int K, bool LAP4, bool DEF, bool FAST, int ARITH = 0>
__global__ void __launch_bounds__(NTHR, 2)
rd_tile_rk(const __grid_constant__ YhK k, const __grid_constant__ TileArgs a) {
  constexpr int H = K;                       // halo (K = 2 or 4: even, keeps pairs 16-byte aligned)
  constexpr int TW = TB + 2 * H;             // tile width / height in cells
  constexpr int NP = TW / 2;                 // pairs per tile row
  constexpr int NPAIR = NP * TW;             // pairs in the tile
  constexpr int SLOTS = (NPAIR + NTHR - 1) / NTHR;
  constexpr int PL = TW * TW;                // doubles per plane
  extern __shared__ __align__(16) double sm[];
  double *s_u0 = sm, *s_v0 = sm + PL, *s_U = sm + 2 * PL, *s_V = sm + 3 * PL;
  double *s_Ju = sm + 4 * PL, *s_Jv = sm + 5 * PL;

  const int tid = threadIdx.x;
  const int nx = k.nx;
  const int gx0 = blockIdx.x * TB - H;                 // global x of tile column 0
  const int ly0 = k.row0 + blockIdx.y * TB - H;        // LOCAL row of tile row 0
  const int dom_lo = -k.jg0, dom_hi = k.nyg - k.jg0;   // local rows that exist globally
  const int out_hi = min(k.row1, k.row0 + (int)(blockIdx.y + 1) * TB);
  const bool gd = FAST ? true : (k.gateDiff != 0);

  double ki[5] = {0, 0, 0, 0, 0}, ws[4] = {0, 0, 0, 0};
  if (K == 4) {
    ki[1] = 0.5; ki[2] = 0.5; ki[3] = 1.0;
    ws[0] = 0.166666666666667; ws[1] = 0.333333333333333; ws[2] = 0.333333333333333; ws[3] = 0.166666666666667;
  } else { ki[1] = 0.5; ws[1] = 1.0; }

  // ---- load u0, v0 and form the stage-0 state and currents --------------------------------
  for (int t = tid; t < NPAIR; t += NTHR) {
    const int ty = t / NP, tx = 2 * (t % NP);
    const int gx = gx0 + tx, ly = ly0 + ty;
    double2 u = make_double2(0, 0), v = u, ju = u, jv = u, U = u, V = u;
    if (gx >= 0 && gx < nx && ly >= dom_lo && ly < dom_hi) {
      const size_t o = (size_t)ly * nx + gx;
      u = *reinterpret_cast<const double2 *>(a.u_in + o);
      v = *reinterpret_cast<const double2 *>(a.v_in + o);
      U.x = u.x + 0.0; U.y = u.y + 0.0; V.x = v.x + 0.0; V.y = v.y + 0.0;   // u0 + (0.0*0.0)
      const int gj = ly + k.jg0;
      ju.x = t_Isum<DEF, FAST>(k, U.x, V.x, FAST ? false : yh_scs(k, gx, gj));
      ju.y = t_Isum<DEF, FAST>(k, U.y, V.y, FAST ? false : yh_scs(k, gx + 1, gj));
      jv.x = t_Iv<DEF>(k, U.x, V.x);
      jv.y = t_Iv<DEF>(k, U.y, V.y);
    }
    const int c = ty * TW + tx;
    *reinterpret_cast<double2 *>(s_u0 + c) = u;
    *reinterpret_cast<double2 *>(s_v0 + c) = v;
    *reinterpret_cast<double2 *>(s_U + c) = U;
    *reinterpret_cast<double2 *>(s_V + c) = V;
    *reinterpret_cast<double2 *>(s_Ju + c) = ju;
    *reinterpret_cast<double2 *>(s_Jv + c) = jv;
  }
  __syncthreads();

  const double q4 = k.qx4 + k.qy4, m2q = -2.0 * q4, mrs2 = -k.rscale * 2.0, rsq = k.rscale * q4;
  const double mrs2q4 = mrs2 * q4;   // the product the reference forms first (left to right, :235)
  double2 ru[SLOTS], rv[SLOTS], du[SLOTS], dv[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; s++) ru[s] = rv[s] = du[s] = dv[s] = make_double2(0.0, 0.0);

  // Geometry of this thread's pairs, once per launch instead of once per stage and phase (the divisions and
  // range tests were 40 % of the executed instructions, ncu profiles/r2a_tile_rk_512_mix.txt): cell index,
  // offsets of the S / N rows (no-flux mirror = index selection inside the tile, helper_functions.cu:69-79),
  // and a bit per stage: "du of this stage is needed here" (tile minus st+1 outer rings); bit K: output row.
  int g_c[SLOTS], g_s[SLOTS], g_n[SLOTS];
  unsigned g_m[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; s++) {
    const int t = tid + s * NTHR;
    g_c[s] = 0; g_s[s] = 0; g_n[s] = 0; g_m[s] = 0;
    if (t >= NPAIR) continue;
    const int ty = t / NP, tx = 2 * (t % NP);
    const int gx = gx0 + tx, ly = ly0 + ty;
    const bool in_dom = gx >= 0 && gx < nx && ly >= dom_lo && ly < dom_hi;
    g_c[s] = ty * TW + tx;
    g_s[s] = (ly - 1 < dom_lo) ? TW : -TW;
    g_n[s] = (ly + 1 >= dom_hi) ? -TW : TW;
    unsigned m = 0;
#pragma unroll
    for (int st = 0; st < K; st++) m |= (in_dom && pair_in_ring(tx, ty, st + 1, TW)) ? (1u << st) : 0u;
    if (ly >= k.row0 && ly < out_hi) m |= 1u << K;
    // tile-edge pairs (tx == 0 / TW-2) hold one cell outside the ring: its value is never read by a valid
    // cell, it only must not read outside the tile
    if ((gx == 0) || (tx == 0)) m |= 1u << (K + 1);
    if ((gx + 2 == nx) || (tx == TW - 2)) m |= 1u << (K + 2);
    g_m[s] = m;
  }

#pragma unroll
  for (int st = 0; st < K; st++) {
    // cells whose du_st is needed: tile minus st+1 outer rings (bit st of g_m)
    // ---- phase A: du, dv of stage st into registers ----------------------------------------
#pragma unroll
    for (int s = 0; s < SLOTS; s++) {
      if (!((g_m[s] >> st) & 1u)) continue;
      const int c = g_c[s], cs = c + g_s[s], cn = c + g_n[s];
      const bool le = (g_m[s] >> (K + 1)) & 1u, re = (g_m[s] >> (K + 2)) & 1u;
      double d[2][2];
#pragma unroll
      for (int f = 0; f < 2; f++) {
        const double *P = f ? s_V : s_U, *J = f ? s_Jv : s_Ju;
        const double2 C = *reinterpret_cast<const double2 *>(P + c);
        const double2 S = *reinterpret_cast<const double2 *>(P + cs);
        const double2 N = *reinterpret_cast<const double2 *>(P + cn);
        const double Wv = le ? C.y : P[c - 1], Ev = re ? C.x : P[c + 2];
        double d0, d1;
        if (ARITH == 1) {   // FAST flavour: collected coefficients, FMA chains (FAST switches: gateDiff on)
          const double2 Jc = *reinterpret_cast<const double2 *>(J + c);
          double q0 = 0.0, q1 = 0.0, jwe0 = 0.0, jwe1 = 0.0, jns0 = 0.0, jns1 = 0.0;
          const double ns0 = N.x + S.x, ns1 = N.y + S.y;
          if (LAP4) {
            const double SWv = le ? S.y : P[cs - 1], SEv = re ? S.x : P[cs + 2];
            const double NWv = le ? N.y : P[cn - 1], NEv = re ? N.x : P[cn + 2];
            const double2 Js = *reinterpret_cast<const double2 *>(J + cs);
            const double2 Jn = *reinterpret_cast<const double2 *>(J + cn);
            const double JW = le ? Jc.y : J[c - 1], JE = re ? Jc.x : J[c + 2];
            q0 = (SWv + NWv) + ns1; q1 = ns0 + (SEv + NEv);     // corners of a cell = (N+S) of its two x neighbours
            jwe0 = JW + Jc.y; jwe1 = Jc.x + JE;
            jns0 = Jn.x + Js.x; jns1 = Jn.y + Js.y;
          }
          if (f == 0) {
            d[f][0] = t_cell_fast<0, LAP4>(a, C.x, Wv + C.y, ns0, q0, Jc.x, jwe0, jns0);
            d[f][1] = t_cell_fast<0, LAP4>(a, C.y, C.x + Ev, ns1, q1, Jc.y, jwe1, jns1);
          } else {
            d[f][0] = t_cell_fast<1, LAP4>(a, C.x, Wv + C.y, ns0, q0, Jc.x, jwe0, jns0);
            d[f][1] = t_cell_fast<1, LAP4>(a, C.y, C.x + Ev, ns1, q1, Jc.y, jwe1, jns1);
          }
          continue;
        }
        if (f == 0) {
          d0 = ((fma(-2.0, C.x, Wv) + C.y) * k.rx + (fma(-2.0, C.x, N.x) + S.x) * k.ry);
          d1 = ((fma(-2.0, C.y, C.x) + Ev) * k.rx + (fma(-2.0, C.y, N.y) + S.y) * k.ry);
        } else if (gd) {
          d0 = ((fma(-2.0, C.x, Wv) + C.y) * k.rx * k.rscale + (fma(-2.0, C.x, N.x) + S.x) * k.ry * k.rscale);
          d1 = ((fma(-2.0, C.y, C.x) + Ev) * k.rx * k.rscale + (fma(-2.0, C.y, N.y) + S.y) * k.ry * k.rscale);
        } else { d0 = 0.0; d1 = 0.0; }
        const double2 Jc = *reinterpret_cast<const double2 *>(J + c);
        if (LAP4 && (f == 0 || gd)) {
          const double SWv = le ? S.y : P[cs - 1], SEv = re ? S.x : P[cs + 2];
          const double NWv = le ? N.y : P[cn - 1], NEv = re ? N.x : P[cn + 2];
          const double2 Js = *reinterpret_cast<const double2 *>(J + cs);
          const double2 Jn = *reinterpret_cast<const double2 *>(J + cn);
          const double JW = le ? Jc.y : J[c - 1], JE = re ? Jc.x : J[c + 2];
          if (f == 0) {   // reactionDiffusion.cu:221-229
            d0 += m2q * (+(Wv - C.x + C.y) + (N.x - C.x + S.x));
            d1 += m2q * (+(C.x - C.y + Ev) + (N.y - C.y + S.y));
            d0 += q4 * (SWv + S.y + NWv + N.y);
            d1 += q4 * (S.x + SEv + N.x + NEv);
          } else {        // :235-239
            d0 += mrs2q4 * (+(Wv - C.x + C.y) + (N.x - C.x + S.x));
            d1 += mrs2q4 * (+(C.x - C.y + Ev) + (N.y - C.y + S.y));
            d0 += rsq * (SWv + S.y + NWv + N.y);
            d1 += rsq * (S.x + SEv + N.x + NEv);
          }
          d0 -= ((fma(-2.0, Jc.x, JW) + Jc.y) * k.fx4 + (fma(-2.0, Jc.x, Jn.x) + Js.x) * k.fy4);
          d1 -= ((fma(-2.0, Jc.y, Jc.x) + JE) * k.fx4 + (fma(-2.0, Jc.y, Jn.y) + Js.y) * k.fy4);
        }
        d[f][0] = d0 - k.dt * Jc.x;   // :498-499
        d[f][1] = d1 - k.dt * Jc.y;
      }
      du[s] = make_double2(d[0][0], d[0][1]);
      dv[s] = make_double2(d[1][0], d[1][1]);
      if (ARITH == 1) {
        ru[s].x = fma(ws[st], du[s].x, ru[s].x); ru[s].y = fma(ws[st], du[s].y, ru[s].y);
        rv[s].x = fma(ws[st], dv[s].x, rv[s].x); rv[s].y = fma(ws[st], dv[s].y, rv[s].y);
      } else {
        ru[s].x += (ws[st] * du[s].x); ru[s].y += (ws[st] * du[s].y);   // :502-503
        rv[s].x += (ws[st] * dv[s].x); rv[s].y += (ws[st] * dv[s].y);
      }
    }
    __syncthreads();   // every read of the stage-st state is done
    // ---- phase B: stage st+1 state and currents, or the final update -----------------------
#pragma unroll
    for (int s = 0; s < SLOTS; s++) {
      if (!((g_m[s] >> st) & 1u)) continue;
      const int c = g_c[s];
      const double2 u0 = *reinterpret_cast<const double2 *>(s_u0 + c);
      const double2 v0 = *reinterpret_cast<const double2 *>(s_v0 + c);
      if (st < K - 1) {
        double2 U, V, ju, jv;
        if (ARITH == 1) {
          U.x = fma(ki[st + 1], du[s].x, u0.x); U.y = fma(ki[st + 1], du[s].y, u0.y);
          V.x = fma(ki[st + 1], dv[s].x, v0.x); V.y = fma(ki[st + 1], dv[s].y, v0.y);
          t_currents_fast<DEF>(k, a, U.x, V.x, ju.x, jv.x);
          t_currents_fast<DEF>(k, a, U.y, V.y, ju.y, jv.y);
        } else {
          U.x = u0.x + (ki[st + 1] * du[s].x); U.y = u0.y + (ki[st + 1] * du[s].y);   // :117-118
          V.x = v0.x + (ki[st + 1] * dv[s].x); V.y = v0.y + (ki[st + 1] * dv[s].y);
          bool s0 = false, s1 = false;
          if (!FAST) {   // live stimulus: global coordinates of the pair from its tile index
            const int t = tid + s * NTHR;
            const int gx = gx0 + 2 * (t % NP), gj = ly0 + t / NP + k.jg0;
            s0 = yh_scs(k, gx, gj); s1 = yh_scs(k, gx + 1, gj);
          }
          ju.x = t_Isum<DEF, FAST>(k, U.x, V.x, s0);
          ju.y = t_Isum<DEF, FAST>(k, U.y, V.y, s1);
          jv.x = t_Iv<DEF>(k, U.x, V.x);
          jv.y = t_Iv<DEF>(k, U.y, V.y);
        }
        *reinterpret_cast<double2 *>(s_U + c) = U;
        *reinterpret_cast<double2 *>(s_V + c) = V;
        *reinterpret_cast<double2 *>(s_Ju + c) = ju;
        *reinterpret_cast<double2 *>(s_Jv + c) = jv;
      } else if ((g_m[s] >> K) & 1u) {   // ring == H here: exactly the output tile
        double2 uo, vo;   // :512-513
        if (ARITH == 1) {
          uo.x = fma(k.tc, ru[s].x, u0.x); uo.y = fma(k.tc, ru[s].y, u0.y);
          vo.x = fma(k.tc, rv[s].x, v0.x); vo.y = fma(k.tc, rv[s].y, v0.y);
        } else {
          uo.x = u0.x + k.tc * ru[s].x; uo.y = u0.y + k.tc * ru[s].y;
          vo.x = v0.x + k.tc * rv[s].x; vo.y = v0.y + k.tc * rv[s].y;
        }
        const int t = tid + s * NTHR;
        const size_t o = (size_t)(ly0 + t / NP) * nx + (gx0 + 2 * (t % NP));
        *reinterpret_cast<double2 *>(a.u_out + o) = uo;
        *reinterpret_cast<double2 *>(a.v_out + o) = vo;
        if (a.vtu && gd) {   // :551-552
          *reinterpret_cast<double2 *>(a.vtu + o) = make_double2(ru[s].x / k.dt, ru[s].y / k.dt);
          *reinterpret_cast<double2 *>(a.vtv + o) = make_double2(rv[s].x / k.dt, rv[s].y / k.dt);
        }
      }
    }
    if (st < K - 1) __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Euler tile kernel: T steps per launch.  ONE pair per thread for the whole launch: the pair's own
// values stay in registers from step to step, only its neighbours come from the ping-pong planes
// in shared memory; index arithmetic and boundary selections are done once.  ~800 threads per
// CTA, two CTAs per SM: the small sheet becomes enough warps to hide the FP64 latency.
// ------------------------------------------------------------------------------------------
template <int T>
struct EulerTile {
  static constexpr int H = (T + 1) & ~1;
  static constexpr int TW = TB + 2 * H;
  static constexpr int NP = TW / 2;
  static constexpr int NPAIR = NP * TW;
  static constexpr int NT = (NPAIR + 31) & ~31;
  static constexpr int PL = TW * TW;
};

template <int T, bool DEF>
__global__ void __launch_bounds__(EulerTile<T>::NT, 2)
rd_tile_euler(const __grid_constant__ YhK k, const __grid_constant__ TileArgs a) {
  using E = EulerTile<T>;
  constexpr int H = E::H, TW = E::TW, NP = E::NP, NPAIR = E::NPAIR, PL = E::PL;
  extern __shared__ __align__(16) double sm[];

  const int tid = threadIdx.x;
  const int nx = k.nx;
  const int ty = tid / NP, tx = 2 * (tid % NP);
  const int gx = blockIdx.x * TB - H + tx;
  const int ly = k.row0 + blockIdx.y * TB - H + ty;        // LOCAL row
  const int dom_lo = -k.jg0, dom_hi = k.nyg - k.jg0;       // local rows that exist globally
  const int out_hi = min(k.row1, k.row0 + (int)(blockIdx.y + 1) * TB);
  const bool in_dom = tid < NPAIR && gx >= 0 && gx < nx && ly >= dom_lo && ly < dom_hi;
  const size_t o = (size_t)blockIdx.z * (size_t)a.sim_stride + (size_t)(in_dom ? ly : 0) * nx + (in_dom ? gx : 0);

  // no-flux mirror = index selection inside the tile (helper_functions.cu:69-79); tile-edge pairs
  // hold one cell outside every ring: it only must not read outside the tile
  const int c = ty * TW + tx;
  const int cs = (ly - 1 < dom_lo) ? c + TW : c - TW;
  const int cn = (ly + 1 >= dom_hi) ? c - TW : c + TW;
  const bool le = (gx == 0) || (tx == 0), re = (gx + 2 == nx) || (tx == TW - 2);

  double2 uC = make_double2(0.0, 0.0), vC = uC;
  if (in_dom) {
    uC = *reinterpret_cast<const double2 *>(a.u_in + o);
    vC = *reinterpret_cast<const double2 *>(a.v_in + o);
    uC.x += 0.0; uC.y += 0.0; vC.x += 0.0; vC.y += 0.0;   // the reference's u0 + (0.0*0.0)
  }
  if (tid < NPAIR) {
    *reinterpret_cast<double2 *>(sm + c) = uC;
    *reinterpret_cast<double2 *>(sm + PL + c) = vC;
  }
  __syncthreads();
  // the thread that OWNS the electrode cell (inside this CTA's output tile) records it at every level
  const bool probe = a.trace != nullptr && in_dom && tx >= H && tx < H + TB && ty >= H && ty < H + TB &&
                     (ly + k.jg0) == k.py && (k.px == gx || k.px == gx + 1);
  double *probe_out = nullptr;
  if (probe) probe_out = a.trace + 2 * ((size_t)(*a.slot) * gridDim.z + blockIdx.z);

#pragma unroll
  for (int s = 1; s <= T; s++) {
    if (probe) {   // state after s-1 steps = the sample taken BEFORE step s (one-step lag of main.cu:1040)
      double *o2 = probe_out + 2 * (size_t)(s - 1) * gridDim.z;
      o2[0] = (k.px == gx) ? uC.x : uC.y;
      o2[1] = (k.px == gx) ? vC.x : vC.y;
    }
    const double *iu = sm + ((s - 1) & 1) * 2 * PL, *iv = iu + PL;
    double *ou = sm + (s & 1) * 2 * PL, *ov = ou + PL;
    const int ring = (H - T) + s;   // valid region shrinks by one ring per step
    bool stim_on = k.stim != 0;
    if (a.period) {
      const int per = a.period[blockIdx.z];
      stim_on = per > 0 && ((a.count0 + s - 1) % per) <= a.duration;
    }
    if (in_dom && pair_in_ring(tx, ty, ring, TW)) {
      const double2 uS = *reinterpret_cast<const double2 *>(iu + cs), uN = *reinterpret_cast<const double2 *>(iu + cn);
      const double uw = le ? uC.y : iu[c - 1], ue = re ? uC.x : iu[c + 2];
      double2 un, vn;
      double du0 = ((fma(-2.0, uC.x, uw) + uC.y) * k.rx + (fma(-2.0, uC.x, uN.x) + uS.x) * k.ry);
      double du1 = ((fma(-2.0, uC.y, uC.x) + ue) * k.rx + (fma(-2.0, uC.y, uN.y) + uS.y) * k.ry);
      double dv0 = 0.0, dv1 = 0.0;
      if (DEF || k.gateDiff) {   // DEF includes gateDiff == 1: no branch inside the step (see rd_fast.cu)
        const double2 vS = *reinterpret_cast<const double2 *>(iv + cs), vN = *reinterpret_cast<const double2 *>(iv + cn);
        const double vw = le ? vC.y : iv[c - 1], ve = re ? vC.x : iv[c + 2];
        dv0 = ((fma(-2.0, vC.x, vw) + vC.y) * k.rx * k.rscale + (fma(-2.0, vC.x, vN.x) + vS.x) * k.ry * k.rscale);
        dv1 = ((fma(-2.0, vC.y, vC.x) + ve) * k.rx * k.rscale + (fma(-2.0, vC.y, vN.y) + vS.y) * k.ry * k.rscale);
      }
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const double u = q ? uC.y : uC.x, v = q ? vC.y : vC.x;
        double du = q ? du1 : du0, dv = q ? dv1 : dv0;
        const bool scs = stim_on && yh_scs_on(k, gx + q, ly + k.jg0);
        // same rewrites as rd_fast.cu::euler_cell (bit-exact identities, see there)
        const double mu_u = DEF ? u : k.mu * u;
        const double X = mu_u * (1.0 - u) * (u - k.alpha) - u * v;
        const double ug = DEF ? u : k.delta * (u - k.gamma);
        const double yv = ug * (k.beta - u) - v;
        const double Y = k.eps * (DEF ? yv : yv - k.theta);
        if (!scs) du = du + k.dt * X;
        else { const double I_sum = -X - 24.7; du = du - k.dt * I_sum; }
        dv = dv + k.dt * Y;
        const double u1 = u + (DEF ? du : k.tc * du), v1 = v + (DEF ? dv : k.tc * dv);
        if (q) { un.y = u1; vn.y = v1; } else { un.x = u1; vn.x = v1; }
      }
      uC = un; vC = vn;
      if (s < T) {
        *reinterpret_cast<double2 *>(ou + c) = uC;
        *reinterpret_cast<double2 *>(ov + c) = vC;
      } else if (ly >= k.row0 && ly < out_hi) {   // ring == H here: exactly the output tile
        *reinterpret_cast<double2 *>(a.u_out + o) = uC;
        *reinterpret_cast<double2 *>(a.v_out + o) = vC;
      }
    }
    if (s < T) __syncthreads();
  }
}

template <typename KernelT>
int set_smem(KernelT kern, size_t smem) {
  YH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return YH_OK;
}

template <int K, bool LAP4, bool DEF, bool FAST, int ARITH = 0>
int launch_rk(const YhK &k, const TileArgs &a, cudaStream_t st) {
  constexpr int TW = TB + 2 * K;
  const size_t smem = (size_t)6 * TW * TW * sizeof(double);
  static bool done[64] = {false};
  int dev = 0;
  YH_CUDA(cudaGetDevice(&dev));
  auto kfn = rd_tile_rk<K, LAP4, DEF, FAST, ARITH>;
  if (!done[dev & 63]) { int rc = set_smem(kfn, smem); if (rc) return rc; done[dev & 63] = true; }
  dim3 grd((k.nx + TB - 1) / TB, (k.row1 - k.row0 + TB - 1) / TB);
  YH_LAUNCH(kfn, grd, NTHR, smem, st, k, a);
  return YH_OK;
}

template <int T, bool DEF>
int launch_euler(const YhK &k, const TileArgs &a, int nsims, cudaStream_t st) {
  constexpr int TW = EulerTile<T>::TW;
  const size_t smem = (size_t)4 * TW * TW * sizeof(double);
  static bool done[64] = {false};
  int dev = 0;
  YH_CUDA(cudaGetDevice(&dev));
  if (!done[dev & 63]) { int rc = set_smem(rd_tile_euler<T, DEF>, smem); if (rc) return rc; done[dev & 63] = true; }
  dim3 grd((k.nx + TB - 1) / TB, (k.row1 - k.row0 + TB - 1) / TB, nsims);
  auto kfn = rd_tile_euler<T, DEF>;
  YH_LAUNCH(kfn, grd, EulerTile<T>::NT, smem, st, k, a);
  return YH_OK;
}

bool is_def(const YhK &k) {
  return (k.mu == 1.0) && (k.delta == 1.0) && (k.gamma == 0.0) && (k.theta == 0.0);
}

}  // namespace

// Policy: tiles win while the sheet cannot fill the streaming pipelines.  Measured (B200, Gcell/s,
// tiles vs strips): 512^2 Euler T=4 102 vs 85, RK4+lap4 14.2 vs 12.8; 1024^2 Euler 158 vs 181,
// RK4+lap4 18.9 vs 20.6 -- the crossover sits between them.  YH_RD_PATH = tile | stream overrides
// (tests run every case through both).
int yh_rd_prefer_tile(long long cells) {
  const char *f = getenv("YH_RD_PATH");
  if (f && f[0] == 't') return 1;
  if (f && f[0] == 's') return 0;
  return cells <= (640ll << 10);
}

int yh_rd_tile_rk_supported(const YhK &k) {
  if (k.timeIntOrder != 2 && k.timeIntOrder != 4) return 0;
  if (!k.neumannBC || k.solidSwitch || k.anisotropy) return 0;
  if ((k.nx & 1) || k.nx < 8) return 0;
  return 1;
}

int yh_launch_rd_tile_rk(const YhK &k, const double *u_in, const double *v_in, double *u_out,
                         double *v_out, double *vtu, double *vtv, cudaStream_t st) {
  if (!yh_rd_tile_rk_supported(k)) return YH_ERR_UNSUPPORTED;
  if (k.row1 <= k.row0) return YH_OK;
  // default switches (gate diffusion on, live stimulus off): the marching kernel, one tile per SM (rd_tile_march.cu)
  if (yh_rd_tile_march_supported(k)) return yh_launch_rd_tile_march(k, u_in, v_in, u_out, v_out, vtu, vtv, st);
  TileArgs a{u_in, v_in, u_out, v_out, vtu, vtv, 0, nullptr, 0, 0, nullptr, nullptr};
  const bool lap4 = k.lap4 != 0, def = is_def(k);
#define YH_T(KK, L) (def ? launch_rk<KK, L, true, false>(k, a, st) : launch_rk<KK, L, false, false>(k, a, st))
  if (k.timeIntOrder == 4 && lap4 && k.gateDiff && !k.stim) {   // the reference's default mode
    if (yh_arithmetic() == YH_ARITH_FAST) {
      const double q = k.qx4 + k.qy4;
      a.uH = k.rx - 2.0 * q; a.uV = k.ry - 2.0 * q; a.uC = -2.0 * (k.rx + k.ry) + 4.0 * q; a.uQ = q;
      a.vH = k.rscale * a.uH; a.vV = k.rscale * a.uV; a.vC = k.rscale * a.uC; a.vQ = k.rscale * a.uQ;
      a.jX = k.fx4; a.jY = k.fy4; a.jC = 2.0 * (k.fx4 + k.fy4) - k.dt; a.neg_eps = -k.eps;
      return def ? launch_rk<4, true, true, true, 1>(k, a, st) : launch_rk<4, true, false, true, 1>(k, a, st);
    }
    return def ? launch_rk<4, true, true, true>(k, a, st) : launch_rk<4, true, false, true>(k, a, st);
  }
  if (k.timeIntOrder == 4) return lap4 ? YH_T(4, true) : YH_T(4, false);
  return lap4 ? YH_T(2, true) : YH_T(2, false);
#undef YH_T
}

__global__ void slot_bump_kernel(unsigned long long *slot, int n) { *slot += (unsigned long long)n; }

int yh_slot_bump(unsigned long long *slot, int n, cudaStream_t st) {
  slot_bump_kernel<<<1, 1, 0, st>>>(slot, n);
  YH_LAUNCH_CHECK();
  return YH_OK;
}

int yh_launch_rd_tile_euler(const YhK &k, int tb, const double *u_in, const double *v_in, double *u_out,
                            double *v_out, int nsims, long long sim_stride, const int *period_d,
                            int duration_it, int count0, cudaStream_t st, double *trace,
                            const unsigned long long *slot) {
  if (!yh_rd_fast_supported(k, tb)) return YH_ERR_UNSUPPORTED;
  if (k.row1 <= k.row0) return YH_OK;
  TileArgs a{u_in, v_in, u_out, v_out, nullptr, nullptr, sim_stride, period_d, duration_it, count0, trace, slot};
  const bool def = is_def(k) && k.tc == 1.0 && k.gateDiff != 0;
#define YH_E(TT) (def ? launch_euler<TT, true>(k, a, nsims, st) : launch_euler<TT, false>(k, a, nsims, st))
  switch (tb) {
    case 1: return YH_E(1);
    case 2: return YH_E(2);
    default: return YH_E(4);
  }
#undef YH_E
}
