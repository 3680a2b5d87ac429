// rd_tile_march.cu -- the reference's DEFAULT step (RK4 + 4th-order Laplacian, gate diffusion on,
// reactionDiffusion.cu:71-93,115-247,498-561) for SMALL sheets (512 x 512 is the reference's own
// default, saveFiles.cu:137-138), second generation of rd_tile.cu::rd_tile_rk.
//
// What the first tile kernel was bound by (ncu, profiles/r2c_tile_rk_512_raw.csv): the shared-memory
// pipe -- 2.5 wavefronts per cell and stage (every pair re-read its 3 x 3 neighbourhood of U, V and the
// 5 points of the two currents, 27 % of the loads bank-conflicted on the 40-double row pitch), LSU 82 %
// busy while active, short-scoreboard the top stall; halving the FP64 work changed nothing.  And 256
// tiles on 148 SMs: the busiest SMs carry two tiles, the others one.
//
// Here: ONE tile per SM (the sheet is cut into <= 148 tiles, strips x bands), a WARP owns R rows of the
// tile's 64 columns (lane = one pair of columns) and MARCHES down them with the S / C / N rows of the stage
// state in registers, so a row is read once per stage (+ 2 halo rows per R); even and odd columns live
// in separate planes: every access is a conflict-free LDS.64 / STS.64 (own pair: E[p], O[p]; west / east
// neighbour: O[p-1], E[p+1]).  0.9 wavefronts per cell and stage.  du / rhs / u0 / v0 of a cell stay in
// registers over the stages (fixed cell -> thread map); stage k+1 values go to the other half of a
// ping-pong, ONE __syncthreads per stage.  No-flux mirrors cost nothing in the stage body: the column
// mirror is a pad column written by the producer (one extra predicated STS in the edge lane), the row
// mirror is an address adjustment of the row load.
//
// Same expressions, same bits as rd_tile.cu / rd_rk.cu / the plain-C oracle in the EXACT flavour.
//
// Code size matters here: the stage over R rows is unrolled (~1.4 K instructions, two copies: with / without the next
// stage state), eight warps run through it a few hundred cycles apart, and every further copy measured SLOWER than
// the work it saved -- rows dealt 6/5 to the warps so that the four schedulers carry 11 instead of 12 rows (a second
// pair of instantiations: 13.1 -> 14.3 us), a plain-stencil copy for all-tissue warps of the masked kernel
// (14.0 -> 15.1 us at 512^2).  Both were dropped (profiles/r2d_tile_march_probe.txt).
#include <stdlib.h>

#include "yh_common.cuh"

namespace {

constexpr int LANES = 32;              // pairs per tile row: the tile is 64 columns wide, halo included
constexpr int ROWS_MAX = 48;           // tile rows a CTA can hold (band + 2 halos)

struct MarchArgs {
  const double *u_in, *v_in;
  double *u_out, *v_out, *vtu, *vtv;
  const uint8_t *solid;                // SOLID variants: 1 = tissue (main.cu:676-680), local rows
  int sw, bh;                          // strip width (outputs per tile row, even), band height
  // FAST arithmetic flavour (yh_set_arithmetic; coefficients as in rd_rkq.cu / rd_tile.cu):
  //   d = cC*C + cH*(W+E) + cV*(N+S) + cQ*(SW+SE+NW+NE) + jC*Jc - jX*(JW+JE) - jY*(JN+JS)
  double uC, uH, uV, uQ, vC, vH, vV, vQ, jC, jX, jY, neg_eps;
};

struct Row4 { double w, x, y, e; };    // west neighbour, own even column, own odd column, east neighbour

__device__ __forceinline__ double m_flip(double t) {
  return __hiloint2double(__double2hiint(t) ^ (int)0x80000000, __double2loint(t));
}
// stimulus off: the negation is the last operation before the value is stored, done on the sign bit
template <bool DEF>
__device__ __forceinline__ double m_Isum(const YhK &k, double u, double v) {
  const double mu_u = DEF ? u : k.mu * u;
  return m_flip(mu_u * (1.0 - u) * (u - k.alpha) - u * v);
}
template <bool DEF>
__device__ __forceinline__ double m_Iv(const YhK &k, double u, double v) {
  const double ug = DEF ? u : k.delta * (u - k.gamma);
  const double yv = ug * (k.beta - u) - v;
  return -(k.eps * (DEF ? yv : yv - k.theta));
}
template <bool DEF>
__device__ __forceinline__ void m_currents_fast(const YhK &k, const MarchArgs &a, double u, double v, double &ju, double &jv) {
  const double mu_u = DEF ? u : k.mu * u;
  const double ug = DEF ? u : k.delta * (u - k.gamma);
  ju = m_flip(fma(mu_u * (1.0 - u), u - k.alpha, -(u * v)));
  const double yv = fma(ug, k.beta - u, -v);
  jv = a.neg_eps * (DEF ? yv : yv - k.theta);
}

// Shared-memory geometry: half-plane = (ROWS_MAX + 2) rows of 32 doubles (one pad row above and below the
// tile), 16 half-planes = {ping, pong} x {U, V, Ju, Jv} x {even, odd}; 32 doubles of padding in front and
// behind, so the west read of lane 0 / the east read of lane 31 stay inside the allocation (their values are
// never used by a valid cell).
constexpr int HP = (ROWS_MAX + 2) * LANES;
constexpr int F_U = 0, F_V = 1, F_JU = 2, F_JV = 3;
__host__ __device__ constexpr int plane(int f, int par) { return (f * 2 + par) * HP; }
constexpr size_t SMEM_BYTES = (size_t)(16 * HP + 2 * LANES) * sizeof(double);

template <bool FULL>
__device__ __forceinline__ Row4 ld_row(const double *p, int f) {
  Row4 r;
  r.x = p[plane(f, 0)];
  r.y = p[plane(f, 1)];
  if (FULL) { r.w = p[plane(f, 1) - 1]; r.e = p[plane(f, 0) + 1]; } else { r.w = 0.0; r.e = 0.0; }
  return r;
}

__device__ __forceinline__ void st_field(double *p, int f, double x, double y, bool on, bool padL, bool padR) {
  if (on) { p[plane(f, 0)] = x; p[plane(f, 1)] = y; }
  if (padL) p[plane(f, 1) - 1] = y;    // column -1 := column 1
  if (padR) p[plane(f, 0) + 1] = x;    // column nx := column nx - 2
}

// du of the pair for field F (0 = u, 1 = v), EXACT flavour: the reference's expressions, operation for
// operation (reactionDiffusion.cu:201-247, 498-499), as in rd_tile.cu.
template <int F, bool LAP4>
__device__ __forceinline__ void pair_du_exact(const YhK &k, double m2q, double q4, const Row4 &S, const Row4 &C, const Row4 &N,
                                              const Row4 &Js, const Row4 &Jc, const Row4 &Jn, double &o0, double &o1) {
  double d0, d1;
  if (F == 0) {
    d0 = ((fma(-2.0, C.x, C.w) + C.y) * k.rx + (fma(-2.0, C.x, N.x) + S.x) * k.ry);
    d1 = ((fma(-2.0, C.y, C.x) + C.e) * k.rx + (fma(-2.0, C.y, N.y) + S.y) * k.ry);
  } else {
    d0 = ((fma(-2.0, C.x, C.w) + C.y) * k.rx * k.rscale + (fma(-2.0, C.x, N.x) + S.x) * k.ry * k.rscale);
    d1 = ((fma(-2.0, C.y, C.x) + C.e) * k.rx * k.rscale + (fma(-2.0, C.y, N.y) + S.y) * k.ry * k.rscale);
  }
  if (LAP4) {
    // F == 0: m2q = -2*(qx4+qy4), q4 = qx4+qy4 (:221-229); F == 1: m2q = (-rscale*2)*q4, q4 = rscale*q4 (:235-239)
    d0 += m2q * (+(C.w - C.x + C.y) + (N.x - C.x + S.x));
    d1 += m2q * (+(C.x - C.y + C.e) + (N.y - C.y + S.y));
    d0 += q4 * (S.w + S.y + N.w + N.y);
    d1 += q4 * (S.x + S.e + N.x + N.e);
    d0 -= ((fma(-2.0, Jc.x, Jc.w) + Jc.y) * k.fx4 + (fma(-2.0, Jc.x, Jn.x) + Js.x) * k.fy4);
    d1 -= ((fma(-2.0, Jc.y, Jc.x) + Jc.e) * k.fx4 + (fma(-2.0, Jc.y, Jn.y) + Js.y) * k.fy4);
  }
  o0 = d0 - k.dt * Jc.x;
  o1 = d1 - k.dt * Jc.y;
}

// Obstacle masks (reactionDiffusion.cu:154-184), as in rd_rk.cu: the six stencil coefficients of a cell depend only
// on the mask of the cell and of its four (mirrored) neighbours -- packed once per launch into 13 bits per cell
// (cxx cxy cxz cyx cyy cyz, 2 bits each, values 0 / 1 / 2, and sc) and kept in a register of the thread that owns
// the cell for all stages.  The coefficient triples are exact 0 / 1 / 2 and only (1,2,1), (2,2,0), (0,2,2), (0,0,0)
// occur per axis for a tissue cell; each equals -- bit for bit, for finite fields -- the plain stencil
// fma(-2, c, A) + B on substituted neighbours (table in rd_fast.cu).  The mask branch has no 4th-order terms (:184).
__device__ __forceinline__ unsigned m_solid_code(bool sc, bool sw, bool se, bool sn, bool ss) {
  const unsigned cxx = (sw && se) && (sw && sc) ? 1u : ((sw && sc) ? 2u : 0u);
  const unsigned cxy = sc ? ((sw || se) ? 2u : 0u) : 0u;
  const unsigned cxz = (sw && se) && (sc && se) ? 1u : ((sc && se) ? 2u : 0u);
  const unsigned cyx = (sn && ss) && (sn && sc) ? 1u : ((sn && sc) ? 2u : 0u);
  const unsigned cyy = sc ? ((sn || ss) ? 2u : 0u) : 0u;
  const unsigned cyz = (sn && ss) && (sc && ss) ? 1u : ((sc && ss) ? 2u : 0u);
  return cxx | (cxy << 2) | (cxz << 4) | (cyx << 6) | (cyy << 8) | (cyz << 10) | ((sc ? 1u : 0u) << 12);
}
__device__ __forceinline__ double m_axis(unsigned cL, unsigned cC, double c, double L, double R) {
  const bool both = cL == 1u;
  const double t = (cL == 2u) ? L : R;
  const double r = fma(-2.0, c, both ? L : t + t) + (both ? R : 0.0);
  return cC ? r : 0.0;
}
template <int F>
__device__ __forceinline__ void pair_du_solid(const YhK &k, unsigned codes, const Row4 &S, const Row4 &C, const Row4 &N,
                                              const Row4 &Jc, double &o0, double &o1) {
  const unsigned k0 = codes & 0xFFFFu, k1 = codes >> 16;   // reactionDiffusion.cu:171-180
  const double x0 = m_axis(k0 & 3u, (k0 >> 2) & 3u, C.x, C.w, C.y), y0 = m_axis((k0 >> 6) & 3u, (k0 >> 8) & 3u, C.x, N.x, S.x);
  const double x1 = m_axis(k1 & 3u, (k1 >> 2) & 3u, C.y, C.x, C.e), y1 = m_axis((k1 >> 6) & 3u, (k1 >> 8) & 3u, C.y, N.y, S.y);
  double d0, d1;
  if (F == 0) {
    d0 = (x0 * k.rx + y0 * k.ry);
    d1 = (x1 * k.rx + y1 * k.ry);
  } else {
    d0 = (x0 * k.rx * k.rscale + y0 * k.ry * k.rscale);
    d1 = (x1 * k.rx * k.rscale + y1 * k.ry * k.rscale);
  }
  o0 = d0 - k.dt * Jc.x;
  o1 = d1 - k.dt * Jc.y;
}

template <int F, bool LAP4>
__device__ __forceinline__ void pair_du_fast(const MarchArgs &a, const Row4 &S, const Row4 &C, const Row4 &N, const Row4 &Js,
                                             const Row4 &Jc, const Row4 &Jn, double &o0, double &o1) {
  const double ns0 = N.x + S.x, ns1 = N.y + S.y;
  double d0 = (F == 0 ? a.uC : a.vC) * C.x, d1 = (F == 0 ? a.uC : a.vC) * C.y;
  d0 = fma(F == 0 ? a.uH : a.vH, C.w + C.y, d0);
  d1 = fma(F == 0 ? a.uH : a.vH, C.x + C.e, d1);
  d0 = fma(F == 0 ? a.uV : a.vV, ns0, d0);
  d1 = fma(F == 0 ? a.uV : a.vV, ns1, d1);
  if (LAP4) {
    const double q0 = (S.w + N.w) + ns1, q1 = ns0 + (S.e + N.e);   // corners of a cell = (N+S) of its two x neighbours
    d0 = fma(F == 0 ? a.uQ : a.vQ, q0, d0);
    d1 = fma(F == 0 ? a.uQ : a.vQ, q1, d1);
    d0 = fma(-a.jX, Jc.w + Jc.y, d0);
    d1 = fma(-a.jX, Jc.x + Jc.e, d1);
    d0 = fma(-a.jY, Jn.x + Js.x, d0);
    d1 = fma(-a.jY, Jn.y + Js.y, d1);
  }
  o0 = fma(a.jC, Jc.x, d0);
  o1 = fma(a.jC, Jc.y, d1);
}

struct MarchCtx {
  const double *cur;                   // this thread's row 0 in the half of the ping-pong that holds the stage state
  double *nxt;                         // ... in the half the next stage state goes to
  double wst, kin, q4u, m2qu, q4v, m2qv;
  int jlo, jhi;                        // rows of the warp (0 .. R-1) whose results are kept at this stage
  int jm_lo, jm_hi;                    // rows (-1 .. R) that are global row -1 / nyg: loads redirected to the mirror row
  bool in_x, out_x, padL, padR;
  size_t o;                            // global element offset of the thread's pair in row 0 of the warp
};

// One stage over the R rows of a warp.  ONE basic block: every row is loaded and computed, only the stores
// are predicated (rows outside the stage's ring or outside the domain produce values nobody reads), so ptxas
// can issue the loads of row j+1 under the arithmetic of row j -- with a branch per row the LDS latency
// and the tail of every row's dependency chain were exposed (2 warps per scheduler cannot hide them).
template <int K, bool LAP4, bool DEF, int ARITH, int R, bool SOLID, bool LAST>
__device__ __forceinline__ void march_stage(const YhK &k, const MarchArgs &a, const MarchCtx &c, const double2 (&u0)[R],
                                            const double2 (&v0)[R], double2 (&ru)[R], double2 (&rv)[R],
                                            const unsigned (&codes)[R]) {
  Row4 Su, Cu, Nu, Sv, Cv, Nv, Sju, Cju, Nju, Sjv, Cjv, Njv;
  Sju = Cju = Nju = Sjv = Cjv = Njv = Row4{0.0, 0.0, 0.0, 0.0};
  {
    const double *p = c.cur + (-1 + ((-1 == c.jm_lo) ? 2 : 0)) * LANES;
    Su = ld_row<LAP4>(p, F_U); Sv = ld_row<LAP4>(p, F_V);
    if (LAP4) { Sju = ld_row<false>(p, F_JU); Sjv = ld_row<false>(p, F_JV); }
    p = c.cur + ((0 == c.jm_lo) ? 2 : 0) * LANES;   // (row 0 of the warp can be global row -1 only when it is inactive)
    Cu = ld_row<true>(p, F_U); Cv = ld_row<true>(p, F_V);
    Cju = ld_row<LAP4>(p, F_JU); Cjv = ld_row<LAP4>(p, F_JV);
  }
#pragma unroll
  for (int j = 0; j < R; j++) {
    {
      const int adj = (j + 1 == c.jm_hi) ? -2 : ((j + 1 == c.jm_lo) ? 2 : 0);
      const double *p = c.cur + (j + 1 + adj) * LANES;
      if (j + 1 < R) {
        Nu = ld_row<true>(p, F_U); Nv = ld_row<true>(p, F_V);
        Nju = ld_row<LAP4>(p, F_JU); Njv = ld_row<LAP4>(p, F_JV);
      } else {
        Nu = ld_row<LAP4>(p, F_U); Nv = ld_row<LAP4>(p, F_V);
        if (LAP4) { Nju = ld_row<false>(p, F_JU); Njv = ld_row<false>(p, F_JV); }
      }
    }
    const bool act = j >= c.jlo && j <= c.jhi;
    double du0, du1, dv0, dv1;
    if (ARITH == 1) {
      pair_du_fast<0, LAP4>(a, Su, Cu, Nu, Sju, Cju, Nju, du0, du1);
      pair_du_fast<1, LAP4>(a, Sv, Cv, Nv, Sjv, Cjv, Njv, dv0, dv1);
      ru[j].x = fma(c.wst, du0, ru[j].x); ru[j].y = fma(c.wst, du1, ru[j].y);
      rv[j].x = fma(c.wst, dv0, rv[j].x); rv[j].y = fma(c.wst, dv1, rv[j].y);
    } else {
      if (SOLID) {
        pair_du_solid<0>(k, codes[j], Su, Cu, Nu, Cju, du0, du1);
        pair_du_solid<1>(k, codes[j], Sv, Cv, Nv, Cjv, dv0, dv1);
      } else {
        pair_du_exact<0, LAP4>(k, c.m2qu, c.q4u, Su, Cu, Nu, Sju, Cju, Nju, du0, du1);
        pair_du_exact<1, LAP4>(k, c.m2qv, c.q4v, Sv, Cv, Nv, Sjv, Cjv, Njv, dv0, dv1);
      }
      ru[j].x += (c.wst * du0); ru[j].y += (c.wst * du1);   // :502-503
      rv[j].x += (c.wst * dv0); rv[j].y += (c.wst * dv1);
    }
    if (!LAST) {
      double Ux, Uy, Vx, Vy, jux, juy, jvx, jvy;
      if (ARITH == 1) {
        Ux = fma(c.kin, du0, u0[j].x); Uy = fma(c.kin, du1, u0[j].y);
        Vx = fma(c.kin, dv0, v0[j].x); Vy = fma(c.kin, dv1, v0[j].y);
        m_currents_fast<DEF>(k, a, Ux, Vx, jux, jvx);
        m_currents_fast<DEF>(k, a, Uy, Vy, juy, jvy);
      } else {
        Ux = u0[j].x + (c.kin * du0); Uy = u0[j].y + (c.kin * du1);   // :117-118
        Vx = v0[j].x + (c.kin * dv0); Vy = v0[j].y + (c.kin * dv1);
        jux = m_Isum<DEF>(k, Ux, Vx); juy = m_Isum<DEF>(k, Uy, Vy);
        jvx = m_Iv<DEF>(k, Ux, Vx); jvy = m_Iv<DEF>(k, Uy, Vy);
      }
      const bool stw = act && c.in_x;
      double *p = c.nxt + j * LANES;
      st_field(p, F_U, Ux, Uy, stw, stw && c.padL, stw && c.padR);
      st_field(p, F_V, Vx, Vy, stw, stw && c.padL, stw && c.padR);
      st_field(p, F_JU, jux, juy, stw, LAP4 && stw && c.padL, LAP4 && stw && c.padR);
      st_field(p, F_JV, jvx, jvy, stw, LAP4 && stw && c.padL, LAP4 && stw && c.padR);
    } else if (act && c.out_x) {   // ring == H here: exactly the output rows of the band
      double2 uo, vo;              // :512-513
      if (ARITH == 1) {
        uo.x = fma(k.tc, ru[j].x, u0[j].x); uo.y = fma(k.tc, ru[j].y, u0[j].y);
        vo.x = fma(k.tc, rv[j].x, v0[j].x); vo.y = fma(k.tc, rv[j].y, v0[j].y);
      } else {
        uo.x = u0[j].x + k.tc * ru[j].x; uo.y = u0[j].y + k.tc * ru[j].y;
        vo.x = v0[j].x + k.tc * rv[j].x; vo.y = v0[j].y + k.tc * rv[j].y;
      }
      const bool sc0 = !SOLID || ((codes[j] >> 12) & 1u), sc1 = !SOLID || ((codes[j] >> 28) & 1u);
      if (SOLID) {   // :521-522 masked cells are exactly 0.0
        uo.x = sc0 ? uo.x : 0.0; uo.y = sc1 ? uo.y : 0.0;
        vo.x = sc0 ? vo.x : 0.0; vo.y = sc1 ? vo.y : 0.0;
      }
      const size_t o = c.o + (size_t)j * k.nx;
      *reinterpret_cast<double2 *>(a.u_out + o) = uo;
      *reinterpret_cast<double2 *>(a.v_out + o) = vo;
      if (a.vtu) {   // :551-552, :529-530
        *reinterpret_cast<double2 *>(a.vtu + o) = make_double2(sc0 ? ru[j].x / k.dt : 0.0, sc1 ? ru[j].y / k.dt : 0.0);
        *reinterpret_cast<double2 *>(a.vtv + o) = make_double2(sc0 ? rv[j].x / k.dt : 0.0, sc1 ? rv[j].y / k.dt : 0.0);
      }
    }
    Su = Cu; Cu = Nu; Sv = Cv; Cv = Nv;
    Sju = Cju; Cju = Nju; Sjv = Cjv; Cjv = Njv;
  }
}

// K = 2 | 4 stages; switches of the reference's default mode compiled in (gateDiff on, live stimulus off).
// ARITH: 0 = exact, 1 = fast.  R = rows per warp.
template <int K, bool LAP4, bool DEF, int ARITH, int R, bool SOLID>
__global__ void __launch_bounds__(LANES * (ROWS_MAX / R), 1)
rd_tile_march(const __grid_constant__ YhK k, const __grid_constant__ MarchArgs a) {
  constexpr int H = K;
  extern __shared__ __align__(16) double sm_raw[];
  double *sm = sm_raw + LANES;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nx = k.nx;
  const int x0 = blockIdx.x * a.sw, x1 = min(nx, x0 + a.sw);
  const int y0 = k.row0 + blockIdx.y * a.bh, y1 = min(k.row1, y0 + a.bh);   // LOCAL output rows
  const int twy = (y1 - y0) + 2 * H;                                        // tile rows
  const int gx = x0 - H + 2 * lane;                                         // global column of the pair
  const int ly0 = y0 - H;                                                   // local row of tile row 0
  const int dom_lo = -k.jg0, dom_hi = k.nyg - k.jg0;                        // local rows that exist globally
  const int r0 = warp * R;                                                  // first tile row of this warp
  const bool in_x = gx >= 0 && gx < nx;
  const bool padL = gx == 0 && lane > 0, padR = gx + 2 == nx && lane < LANES - 1;   // the pad column must lie inside the tile row
  const bool out_x = gx >= x0 && gx < x1;
  // tile rows this warp may ever touch: inside the tile and inside the global domain
  const int t_lo = max(0, dom_lo - ly0), t_hi = min(twy - 1, dom_hi - 1 - ly0);
  // row mirror (helper_functions.cu:69-79): global row -1 is row 1, row nyg is row nyg - 2 -- the load of
  // tile row t (j = t - r0 in -1 .. R) is redirected two rows down / up
  const int jm_lo = (dom_lo - 1 - ly0) - r0, jm_hi = (dom_hi - ly0) - r0;

  double *base = sm + (r0 + 1) * LANES + lane;   // tile row r0 of plane (0, 0), this lane

  // ---- u0, v0 of the thread's cells; stage-0 state and currents into the ping half ----------------------
  double2 u0[R], v0[R], ru[R], rv[R];
  unsigned codes[R];
#pragma unroll
  for (int j = 0; j < R; j++) {
    const int t = r0 + j;
    u0[j] = v0[j] = ru[j] = rv[j] = make_double2(0.0, 0.0);
    codes[j] = 0u;
    if (t >= t_lo && t <= t_hi && in_x) {
      const size_t o = (size_t)(ly0 + t) * nx + gx;
      if (SOLID && t >= 1 && t <= twy - 2) {   // ring-0 rows are inputs only: no code needed, no mask row beyond the tile read
        const int ly = ly0 + t, gj = ly + k.jg0;
        const uint8_t *mc = a.solid + (size_t)ly * nx;
        const uint8_t *mS = a.solid + (size_t)(yh_mir(gj - 1, k.nyg) - k.jg0) * nx;   // S = j-1 (:149)
        const uint8_t *mN = a.solid + (size_t)(yh_mir(gj + 1, k.nyg) - k.jg0) * nx;   // N = j+1 (:150)
        const bool c0m = mc[gx] != 0, c1m = mc[gx + 1] != 0;
        const bool w0 = mc[yh_mir(gx - 1, nx)] != 0, e1 = mc[yh_mir(gx + 2, nx)] != 0;
        codes[j] = m_solid_code(c0m, w0, c1m, mN[gx] != 0, mS[gx] != 0) |
                   (m_solid_code(c1m, c0m, e1, mN[gx + 1] != 0, mS[gx + 1] != 0) << 16);
      }
      u0[j] = *reinterpret_cast<const double2 *>(a.u_in + o);
      v0[j] = *reinterpret_cast<const double2 *>(a.v_in + o);
      const double Ux = u0[j].x + 0.0, Uy = u0[j].y + 0.0, Vx = v0[j].x + 0.0, Vy = v0[j].y + 0.0;   // u0 + (0.0*0.0)
      double jux, juy, jvx, jvy;
      if (ARITH == 1) {
        m_currents_fast<DEF>(k, a, Ux, Vx, jux, jvx);
        m_currents_fast<DEF>(k, a, Uy, Vy, juy, jvy);
      } else {
        jux = m_Isum<DEF>(k, Ux, Vx); juy = m_Isum<DEF>(k, Uy, Vy);
        jvx = m_Iv<DEF>(k, Ux, Vx); jvy = m_Iv<DEF>(k, Uy, Vy);
      }
      double *p = base + j * LANES;
      st_field(p, F_U, Ux, Uy, true, padL, padR);
      st_field(p, F_V, Vx, Vy, true, padL, padR);
      st_field(p, F_JU, jux, juy, true, LAP4 && padL, LAP4 && padR);
      st_field(p, F_JV, jvx, jvy, true, LAP4 && padL, LAP4 && padR);
    }
  }
  __syncthreads();

  const double q4u = k.qx4 + k.qy4, m2qu = -2.0 * q4u;
  const double mrs2 = -k.rscale * 2.0, q4v = k.rscale * q4u, m2qv = mrs2 * q4u;   // left to right, :235

#pragma unroll 1
  for (int st = 0; st < K; st++) {
    const double *cur = base + (st & 1) * 8 * HP;
    double *nxt = base + ((st & 1) ^ 1) * 8 * HP;
    // RK weights (reactionDiffusion.cu:71-93): stage state u0 + kin*du, rhs += wst*du
    const double wst = K == 4 ? ((st == 0 || st == 3) ? 0.166666666666667 : 0.333333333333333) : (st == 0 ? 0.0 : 1.0);
    const double kin = K == 4 ? (st == 2 ? 1.0 : 0.5) : 0.5;
    // rows of this warp whose du is needed at this stage: tile minus st+1 outer rings, inside the domain
    const int jlo = max(st + 1, t_lo) - r0, jhi = min(twy - 2 - st, t_hi) - r0;
    if (jlo < R && jhi >= 0 && jlo <= jhi) {
      MarchCtx c{cur, nxt, wst, kin, q4u, m2qu, q4v, m2qv, jlo, jhi, jm_lo, jm_hi, in_x, out_x, padL, padR,
                 (size_t)(ly0 + r0) * nx + gx};
      if (st < K - 1) march_stage<K, LAP4, DEF, ARITH, R, SOLID, false>(k, a, c, u0, v0, ru, rv, codes);
      else march_stage<K, LAP4, DEF, ARITH, R, SOLID, true>(k, a, c, u0, v0, ru, rv, codes);
    }
    if (st < K - 1) __syncthreads();
  }
}

// Cut rows x nx outputs into strips x bands: <= 56 (K = 4) outputs per tile row, band + 2K <= ROWS_MAX rows; the
// cost of a launch is waves x (warp-rows of one tile) -- columns cost the same whether they hold outputs or not.
void pick_tiling(int nx, int rows, int K, int nsm, int *sw_out, int *bh_out) {
  const int sw_max = 2 * LANES - 2 * K, bh_max = ROWS_MAX - 2 * K;
  double best = 1e300;
  *sw_out = sw_max; *bh_out = bh_max;
  const int nsx_min = (nx + sw_max - 1) / sw_max, nsy_min = (rows + bh_max - 1) / bh_max;
  for (int nsx = nsx_min; nsx <= nsx_min + 6 && nsx <= (nx + 1) / 2; nsx++) {
    int sw = (nx + nsx - 1) / nsx;
    sw += sw & 1;
    const int nsx_eff = (nx + sw - 1) / sw;
    for (int nsy = nsy_min; nsy <= nsy_min + 4 * nsm && nsy <= rows; nsy++) {
      const int bh = (rows + nsy - 1) / nsy, nsy_eff = (rows + bh - 1) / bh;
      const long long tiles = (long long)nsx_eff * nsy_eff;
      const long long waves = (tiles + nsm - 1) / nsm;
      double work = 0.0;   // active rows over the stages + the load pass
      for (int s = 0; s < K; s++) work += bh + 2 * K - 2 * (s + 1);
      work += 0.25 * (bh + 2 * K) + 6.0;   // load pass, barriers
      const double cost = (double)waves * work;
      if (cost < best) { best = cost; *sw_out = sw; *bh_out = bh; }
    }
  }
}

// the search costs ~10 us of host time: remember the last answers of this thread (a step loop asks for the same one)
void pick_tiling_memo(int nx, int rows, int stages, int nsm, int *sw_out, int *bh_out) {
  struct Memo { int nx, rows, stages, nsm, sw, bh; };
  static thread_local Memo memo[4] = {};
  static thread_local int memo_next = 0;
  for (const Memo &m : memo)
    if (m.nx == nx && m.rows == rows && m.stages == stages && m.nsm == nsm && m.sw > 0) { *sw_out = m.sw; *bh_out = m.bh; return; }
  pick_tiling(nx, rows, stages, nsm, sw_out, bh_out);
  memo[memo_next] = Memo{nx, rows, stages, nsm, *sw_out, *bh_out};
  memo_next = (memo_next + 1) & 3;
}

template <int K, bool LAP4, bool DEF, int ARITH, int R, bool SOLID = false>
int launch_march(const YhK &k, MarchArgs &a, cudaStream_t st) {
  static bool done[64] = {false};
  static int nsm[64] = {0};
  int dev = 0;
  YH_CUDA(cudaGetDevice(&dev));
  auto kfn = rd_tile_march<K, LAP4, DEF, ARITH, R, SOLID>;
  if (!done[dev & 63]) {
    YH_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    YH_CUDA(cudaDeviceGetAttribute(&nsm[dev & 63], cudaDevAttrMultiProcessorCount, dev));
    done[dev & 63] = true;
  }
  const int rows = k.row1 - k.row0;
  const char *e = getenv("YH_MARCH_TILING");   // "sw,bh" override (tuning)
  if (!(e && sscanf(e, "%d,%d", &a.sw, &a.bh) == 2 && a.sw > 0 && !(a.sw & 1) && a.sw <= 2 * LANES - 2 * K && a.bh > 0 &&
        a.bh <= ROWS_MAX - 2 * K))
    pick_tiling_memo(k.nx, rows, K, nsm[dev & 63], &a.sw, &a.bh);
  const int nseg = (a.bh + 2 * K + R - 1) / R;
  dim3 grd((k.nx + a.sw - 1) / a.sw, (rows + a.bh - 1) / a.bh);
  // (programmatic dependent launch -- griddepcontrol.launch_dependents at the top, .wait before the first global
  // read, cudaLaunchAttributeProgrammaticStreamSerialization -- was measured and dropped: 13.1 = 13.1 us exact,
  // 9.9 vs 10.1 fast; a CTA holds 205 KB of shared memory, so the next grid cannot move in before this one leaves)
  YH_LAUNCH(kfn, grd, LANES * nseg, SMEM_BYTES, st, k, a);
  return YH_OK;
}

}  // namespace

extern "C" int yh_rd_tile_march_tiling(int nx, int rows, int stages, int n_sm, int tiling[2]) {
  if (!tiling || nx < 8 || (nx & 1) || rows < 1 || (stages != 2 && stages != 4) || n_sm < 1) {
    yh_set_error("yh_rd_tile_march_tiling: bad arguments");
    return YH_ERR_INVALID_ARG;
  }
  pick_tiling(nx, rows, stages, n_sm, &tiling[0], &tiling[1]);
  return YH_OK;
}

// solidSwitch: only through yh_launch_rd_tile_march_solid (the caller must hold the mask)
int yh_rd_tile_march_supported(const YhK &k) {
  if (k.solidSwitch) return 0;
  return yh_rd_tile_march_solid_supported(k);
}

int yh_rd_tile_march_solid_supported(const YhK &k) {
  if (k.timeIntOrder != 2 && k.timeIntOrder != 4) return 0;
  if (!k.neumannBC || k.anisotropy) return 0;
  if (!k.gateDiff || k.stim) return 0;
  if ((k.nx & 1) || k.nx < 8 || k.nyg < 4) return 0;
  const char *f = getenv("YH_TILE_RK");   // march | cell (A/B, tests)
  if (f && f[0] == 'c') return 0;
  return 1;
}

// Obstacle masks: the exact flavour only (the mask coefficients are the reference's expressions), RK2 / RK4.
int yh_launch_rd_tile_march_solid(const YhK &k, const double *u_in, const double *v_in, double *u_out, double *v_out,
                                  double *vtu, double *vtv, const uint8_t *solid, cudaStream_t st) {
  if (!k.solidSwitch || !solid || !yh_rd_tile_march_solid_supported(k)) return YH_ERR_UNSUPPORTED;
  if (k.row1 <= k.row0) return YH_OK;
  MarchArgs a{};
  a.u_in = u_in; a.v_in = v_in; a.u_out = u_out; a.v_out = v_out; a.vtu = vtu; a.vtv = vtv; a.solid = solid;
  const bool def = (k.mu == 1.0) && (k.delta == 1.0) && (k.gamma == 0.0) && (k.theta == 0.0);
  if (k.timeIntOrder == 4)
    return def ? launch_march<4, false, true, 0, 6, true>(k, a, st) : launch_march<4, false, false, 0, 6, true>(k, a, st);
  return def ? launch_march<2, false, true, 0, 6, true>(k, a, st) : launch_march<2, false, false, 0, 6, true>(k, a, st);
}

int yh_launch_rd_tile_march(const YhK &k, const double *u_in, const double *v_in, double *u_out, double *v_out,
                            double *vtu, double *vtv, cudaStream_t st) {
  if (!yh_rd_tile_march_supported(k)) return YH_ERR_UNSUPPORTED;
  if (k.row1 <= k.row0) return YH_OK;
  MarchArgs a{};
  a.u_in = u_in; a.v_in = v_in; a.u_out = u_out; a.v_out = v_out; a.vtu = vtu; a.vtv = vtv;
  const bool def = (k.mu == 1.0) && (k.delta == 1.0) && (k.gamma == 0.0) && (k.theta == 0.0);
  const bool lap4 = k.lap4 != 0;
  const int arith = yh_arithmetic() == YH_ARITH_FAST ? 1 : 0;
  if (arith) {
    const double q = lap4 ? k.qx4 + k.qy4 : 0.0;
    a.uH = k.rx - 2.0 * q; a.uV = k.ry - 2.0 * q; a.uC = -2.0 * (k.rx + k.ry) + 4.0 * q; a.uQ = q;
    a.vH = k.rscale * a.uH; a.vV = k.rscale * a.uV; a.vC = k.rscale * a.uC; a.vQ = k.rscale * a.uQ;
    a.jX = k.fx4; a.jY = k.fy4; a.jC = (lap4 ? 2.0 * (k.fx4 + k.fy4) : 0.0) - k.dt; a.neg_eps = -k.eps;
  }
  // rows per warp, measured at 512^2 (B200, us per step, R = 6 | 4): exact 13.3 | 13.5, fast 10.9 | 10.1
  static const int rows_env = [] { const char *e = getenv("YH_MARCH_R"); return e ? atoi(e) : 0; }();   // 4 | 6 (tuning; 3 rows per warp spilled and was no faster)
  const int rows_per_warp = rows_env ? rows_env : (arith ? 4 : 6);
#define YH_M(KK, L, D, A) (rows_per_warp == 4 ? launch_march<KK, L, D, A, 4>(k, a, st) : launch_march<KK, L, D, A, 6>(k, a, st))
#define YH_MD(KK, L, A) (def ? YH_M(KK, L, true, A) : YH_M(KK, L, false, A))
#define YH_ML(KK, A) (lap4 ? YH_MD(KK, true, A) : YH_MD(KK, false, A))
  if (k.timeIntOrder == 4) return arith ? YH_ML(4, 1) : YH_ML(4, 0);
  return arith ? YH_ML(2, 1) : YH_ML(2, 0);
#undef YH_ML
#undef YH_MD
#undef YH_M
}
