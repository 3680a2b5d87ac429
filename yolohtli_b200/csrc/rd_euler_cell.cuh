// rd_euler_cell.cuh -- what the temporally blocked Euler kernels (rd_fast.cu: column pairs per
// thread; rd_quad.cu: four columns per thread) share: launch arguments, the one-cell update with the
// reference's expression order (reactionDiffusion.cu:131-141,154-184,188-197,498-537), the fused APD
// epilogue, cp.async helpers and the chunk-height choice.
#pragma once

#include "yh_common.cuh"

namespace yh_euler {

struct FastArgs {
  const double *u_in, *v_in;
  double *u_out, *v_out;
  int RY;                 // output rows per chunk
  long long sim_stride;   // elements between stacked independent sheets (blockIdx.z)
  const int *period;      // per-sheet pacing period in steps (NULL: k.stim decides)
  int duration, count0;   // stimulus on while (step % period) <= duration; step of level 1
  YhApd apd;              // fused APD bookkeeping (STIM variants only); apd.APD1 == NULL: off
  const uint8_t *pat;     // SOLID variants: 5-bit mask pattern per cell (solid_pattern_kernel)
  // FAST arithmetic flavour (rd_quad.cu, yh_set_arithmetic): the same update with its coefficients collected,
  //   un = uC*u + uH*(W+E) + uV*(N+S) + uT*X,   vn = vC*v + vH*(vW+vE) + vV*(vN+vS) + vT*yv
  double uC, uH, uV, uT, vC, vH, vV, vT;
};

// The APD state machine of one cell (spaceAPD.cu:296-342), entered only when the step crossed
// the 0.15 threshold there: without a crossing neither front nor back changes, so none of the
// branches can fire and sAPD / dAPD keep their values (they must have been initialised by one
// full yh_sapd pass).  Argument roles as the reference calls it after the swap (main.cu:1035):
// "uold" := the NEW state, "unew" := the OLD state.
static __device__ __noinline__ void apd_event(const YhK &k, const YhApd &A, size_t c, double uo, double un, int count) {
  const double apdTh = 0.15;
  const bool sc = A.stimulate ? A.stimArea[c] != 0 : true;
  double fr = A.front[c], bk = A.back[c];
  if ((un > apdTh) && (uo < apdTh) && sc) fr = k.dt * (count - (un - apdTh) / (un - uo));
  if ((un < apdTh) && (uo > apdTh) && sc) bk = k.dt * (count - (un - apdTh) / (un - uo));
  bool first = A.first[c] != 0;
  double apd1 = A.APD1[c], apd2 = A.APD2[c];
  if ((bk > 0.0) && (fr > 0.0) && (first == false) && sc) { apd1 = bk - fr; A.APD1[c] = apd1; fr = 0.0; bk = 0.0; first = true; }
  if ((bk > 0.0) && (fr > 0.0) && first && sc) { apd2 = bk - fr; A.APD2[c] = apd2; fr = 0.0; bk = 0.0; first = false; }
  A.front[c] = fr; A.back[c] = bk; A.first[c] = first ? 1 : 0;
  if (A.stimulate) {
    double s = (apd1 - apd2 > 0.0) && sc ? 1.0 : -1.0;
    s *= (double)sc;
    A.sAPD[c] = s;
    double d = apd2;
    d *= (double)sc;
    A.dAPD[c] = d;
  } else {
    A.sAPD[c] = (apd1 - apd2 > 0.0) ? 1.0 : -1.0;
  }
}

__device__ __forceinline__ void cp_async16(unsigned smem, const void *gmem, bool valid) {
  int sz = valid ? 16 : 0;   // src-size 0 => 16 bytes of zero fill, nothing read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem), "l"(gmem), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async16s(unsigned smem, const void *gmem, int src_size) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem), "l"(gmem), "r"(src_size)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// One Euler update of one cell (reactionDiffusion.cu:131-141,188-197,498-513) for CANONICAL
// inputs (no -0.0).  With X = mu*u*(1-u)*(u-alpha) - u*v the reference computes
// I_sum = -X - (scs ? 24.7 : 0.0) and du -= dt*I_sum; for !scs that is du + dt*X bit for bit
// (negation commutes with IEEE rounding, x - 0.0 == x, a - (-b) == a + b).  Likewise
// dv - dt*(-(eps*Y)) == dv + dt*(eps*Y), rhs = 0.0 + 1.0*du only differs from du in the sign
// of a zero that u + tc*rhs cannot see.
// DEF (the reference's default constants tc = mu = delta = 1, gamma = theta = 0,
// saveFiles.cu:220-227) drops the operations that are exact identities in IEEE arithmetic:
// 1.0*x == x, x - 0.0 == x.  2.0*u is exact, so fma(-2.0, u, w) == w - 2.0*u bit for bit.
// Everything after the Laplacian: ionic terms and the update (du, dv hold the diffusion part).
template <bool DEF>
__device__ __forceinline__ void euler_finish(const YhK &k, double u, double v, double du, double dv,
                                             bool scs, double &un, double &vn) {
  const double mu_u = DEF ? u : k.mu * u;
  const double X = mu_u * (1.0 - u) * (u - k.alpha) - u * v;
  const double ug = DEF ? u : k.delta * (u - k.gamma);
  const double yv = ug * (k.beta - u) - v;
  const double Y = k.eps * (DEF ? yv : yv - k.theta);
  if (!scs) {
    du = du + k.dt * X;
  } else {   // inside the stimulus disc: the literal expression
    const double I_sum = -X - 24.7;
    du = du - k.dt * I_sum;
  }
  dv = dv + k.dt * Y;
  un = u + (DEF ? du : k.tc * du);
  vn = v + (DEF ? dv : k.tc * dv);
}

// FAST flavour of one Euler update (no stimulus, no masks): 19 FP64 instructions instead of 32, FMA chains.
template <bool DEF>
__device__ __forceinline__ void euler_cell_fast(const YhK &k, const FastArgs &a, double u, double v, double uW,
                                                double uE, double uN, double uS, double vW, double vE, double vN,
                                                double vS, double &un, double &vn) {
  const double mu_u = DEF ? u : k.mu * u;
  const double X = fma(mu_u * (1.0 - u), u - k.alpha, -(u * v));
  const double ug = DEF ? u : k.delta * (u - k.gamma);
  double yv = fma(ug, k.beta - u, -v);
  if (!DEF) yv -= k.theta;
  double t = a.uC * u;
  t = fma(a.uH, uW + uE, t);
  t = fma(a.uV, uN + uS, t);
  un = fma(a.uT, X, t);
  double s = a.vC * v;
  s = fma(a.vH, vW + vE, s);
  s = fma(a.vV, vN + vS, s);
  vn = fma(a.vT, yv, s);
}

template <bool DEF>
__device__ __forceinline__ void euler_cell(const YhK &k, double u, double v, double uW, double uE,
                                           double uN, double uS, double vW, double vE, double vN,
                                           double vS, bool scs, double &un, double &vn) {
  const double du = ((fma(-2.0, u, uW) + uE) * k.rx + (fma(-2.0, u, uN) + uS) * k.ry);
  double dv = 0.0;
  // DEF includes gateDiff == 1 (the reference's default, saveFiles.cu:126): a run-time branch here,
  // uniform as it is, cuts the cell into basic blocks that ptxas schedules one by one -- the ncu
  // source page showed each block as a serial DADD/DMUL chain waiting on its own latency
  if (DEF || k.gateDiff)
    dv = ((fma(-2.0, v, vW) + vE) * k.rx * k.rscale + (fma(-2.0, v, vN) + vS) * k.ry * k.rscale);
  euler_finish<DEF>(k, u, v, du, dv, scs, un, vn);
}

// Obstacle masks (reactionDiffusion.cu:154-184, 515-537).  pat = sc | sw<<1 | se<<2 | sn<<3 | ss<<4
// (mask of the cell and of its mirrored W / E / N = j+1 / S = j-1 neighbours).  The reference
// multiplies by coefficient triples that are exact 0 / 1 / 2 (:162-169); per axis, for a tissue
// cell, only four triples occur and each reduces -- bit for bit, for finite fields -- to the plain
// stencil on substituted neighbours, r = fma(-2, u, A) + B:
//     both neighbours tissue  (1,2,1):  1*W - 2*u + 1*E            A = W,    B = E
//     only W                  (2,2,0):  2*W - 2*u + 0*E            A = W+W,  B = 0   (2*W exact;
//     only E                  (0,2,2):  0*W - 2*u + 2*E            A = E+E,  B = 0    one rounding)
//     neither                 (0,0,0):  zero                       r = 0
// (x + (+-0) == x for x != 0, and the sign of a zero du cannot reach the output: rhs = 0.0 + 1.0*du,
// DESIGN.md "zero signs").  Non-tissue cells come out as exactly 0.0.  No multiplies, one DADD per
// axis and field, the rest are selects -- the FP64 pipe is what this kernel is short of.
template <bool DEF>
__device__ __forceinline__ void euler_cell_solid(const YhK &k, unsigned pat, double u, double v,
                                                 double uW, double uE, double uN, double uS,
                                                 double vW, double vE, double vN, double vS, bool scs,
                                                 double &un, double &vn) {
  const bool sc = pat & 1u, sw = pat & 2u, se = pat & 4u, sn = pat & 8u, ss = pat & 16u;
  const bool xb = sw && se, yb = sn && ss, xany = sw || se, yany = sn || ss;
  double t, rx_, ry_;
  t = sw ? uW : uE; rx_ = fma(-2.0, u, xb ? uW : t + t) + (xb ? uE : 0.0);
  t = sn ? uN : uS; ry_ = fma(-2.0, u, yb ? uN : t + t) + (yb ? uS : 0.0);
  const double du = (xany ? rx_ : 0.0) * k.rx + (yany ? ry_ : 0.0) * k.ry;
  double dv = 0.0;
  if (DEF || k.gateDiff) {
    t = sw ? vW : vE; rx_ = fma(-2.0, v, xb ? vW : t + t) + (xb ? vE : 0.0);
    t = sn ? vN : vS; ry_ = fma(-2.0, v, yb ? vN : t + t) + (yb ? vS : 0.0);
    dv = (xany ? rx_ : 0.0) * k.rx * k.rscale + (yany ? ry_ : 0.0) * k.ry * k.rscale;
  }
  euler_finish<DEF>(k, u, v, du, dv, scs, un, vn);
  if (!sc) { un = 0.0; vn = 0.0; }   // :521-522
}

// Chunk height.  Every CTA does the same work, so the grid should fill the machine in whole
// waves: pick the number of row chunks C that maximises
//     (CTAs / (waves * slots))  *  (RY / (RY + 3T))        [wave fill x pipeline-fill loss]
// where slots = 148 SMs x resident CTAs per SM of this kernel variant.
static inline int pick_ry(int rows, int strips, int nsims, int T, int slots) {
  const int min_ry = 4;
  int best_ry = rows;
  double best = -1.0;
  const int cmax = rows / min_ry > 0 ? rows / min_ry : 1;
  for (int C = 1; C <= cmax; C++) {
    const int ry = (rows + C - 1) / C;
    if (ry > 512 && ry > min_ry) continue;   // measured (16384^2): chunks of 256-512 rows beat ~1000-row ones by 1.7 %
    const int chunks = (rows + ry - 1) / ry;
    const long long ctas = (long long)strips * chunks * nsims;
    const long long waves = (ctas + slots - 1) / slots;
    const double fill = (double)ctas / (double)(waves * slots);
    const double eff = fill * (double)ry / (double)(ry + 3 * T);
    if (eff > best + 1e-9) { best = eff; best_ry = ry; }
    if (ry <= min_ry) break;
  }
  return best_ry;
}

// Chunk height of the four-columns-per-thread kernels (rd_quad.cu, rd_rkq.cu).  Measured on B200
// (tools/gpu_ry.sh, 16384 columns): a grid that is ONE full wave is the worst case -- every CTA is in the
// same phase, nothing fills in behind the first CTA that finishes (2048 rows: RY = 512, 0.93 wave:
// 338 Gcell/s; RY = 128, 3.7 waves: 372) -- and the rows of pipeline fill cost less than full rows (not every
// level is active).  Score = RY / (RY + 2*depth) * w / (w + 1/4), w = CTAs / resident slots.
static inline int pick_ry_waves(int rows, int strips, int nsims, int depth, int slots) {
  int best_ry = rows;
  double best = -1.0;
  for (int C = 1; C <= rows; C++) {
    const int ry = (rows + C - 1) / C;
    if (ry > 512) continue;
    if (ry < 16 && C > 1) break;
    const int chunks = (rows + ry - 1) / ry;
    const double w = (double)strips * chunks * nsims / (double)slots;
    const double eff = (double)ry / (double)(ry + 2 * depth) * w / (w + 0.25);
    if (eff > best + 1e-9) { best = eff; best_ry = ry; }
  }
  return best_ry;
}

}  // namespace yh_euler
