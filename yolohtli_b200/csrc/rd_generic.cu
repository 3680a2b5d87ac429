// rd_generic.cu -- every mode of reactionDiffusion_kernel (reactionDiffusion.cu:26-566) with
// SYNCHRONOUS Runge-Kutta stages: one launch per stage, stage state and rhs accumulators in
// library workspace, g_in never written, no J array (J at a neighbour is I_sum / I_v of the
// neighbour's stage state, recomputed pointwise).  This is the all-modes correctness path;
// the HBM-roofline path for the Euler / 5-point modes is rd_fast.cu.
//
// Expression order follows the reference literally; the library is compiled with
// --fmad=false so results are bit-identical to the reference's own kernels (built with the
// same flag) in its race-free modes and to the plain-C oracle in every mode.
#include "yh_common.cuh"

namespace {

struct Stage {
  const double *U, *V;   // stage state (local rows)
  bool add0;             // first stage: state = u0 + (0.0*0.0), formed on the fly
  __device__ __forceinline__ double u(int idx) const { return add0 ? U[idx] + 0.0 : U[idx]; }
  __device__ __forceinline__ double v(int idx) const { return add0 ? V[idx] + 0.0 : V[idx]; }
};

// local linear index of GLOBAL cell (i, gj)
__device__ __forceinline__ int LIDX(const YhK &k, int i, int gj) { return i + k.nx * (gj - k.jg0); }

__device__ __forceinline__ bool solid_at(const YhK &k, const uint8_t *solid, int i, int gj) {
  return (i >= 0 && i < k.nx && gj >= 0 && gj < k.nyg) ? solid[LIDX(k, i, gj)] != 0 : false;
}

// diagonal neighbours with per-axis mirror (reactionDiffusion.cu:203-217), global coords
__device__ __forceinline__ void diag_idx(const YhK &k, int i, int j, int &SW, int &SE, int &NW, int &NE) {
  const int nx = k.nx, ny = k.nyg;
  const int iw = (i > 0) ? i - 1 : i + 1, ie = (i < nx - 1) ? i + 1 : i - 1;
  const int js = (j > 0) ? j - 1 : j + 1, jn = (j < ny - 1) ? j + 1 : j - 1;
  SW = LIDX(k, iw, js); SE = LIDX(k, ie, js); NW = LIDX(k, iw, jn); NE = LIDX(k, ie, jn);
}

template <bool IS_U>
__device__ __forceinline__ double G(const Stage &s, int idx) { return IS_U ? s.u(idx) : s.v(idx); }

// anisotropy boundary corrections (reactionDiffusion.cu:267-307), one field
template <bool IS_U>
__device__ double aniso_term(const YhK &k, const Stage &s, int i, int j, int SW, int SE, int NW,
                             int NE, double &edge) {
  const int nx = k.nx, ny = k.nyg;
  const int c = LIDX(k, i, j);
  const double rbx = k.rbx, rby = k.rby;
  double b_S = (j > 0) ? 0.0 : ((j == 0 && (i == 0 || i == (nx - 1))) ? 0.0
             : rby * (G<IS_U>(s, LIDX(k, i + 1, j)) - G<IS_U>(s, LIDX(k, i - 1, j))));
  double b_N = (j < (ny - 1)) ? 0.0 : ((j == (ny - 1) && (i == 0 || i == (nx - 1))) ? 0.0
             : -rby * (G<IS_U>(s, LIDX(k, i + 1, j)) - G<IS_U>(s, LIDX(k, i - 1, j))));
  double b_W = (i > 0) ? 0.0 : ((i == 0 && (j == 0 || j == (ny - 1))) ? 0.0
             : rbx * (G<IS_U>(s, LIDX(k, i, j + 1)) - G<IS_U>(s, LIDX(k, i, j - 1))));
  double b_E = (i < (nx - 1)) ? 0.0 : ((i == (nx - 1) && (j == 0 || j == (ny - 1))) ? 0.0
             : -rbx * (G<IS_U>(s, LIDX(k, i, j + 1)) - G<IS_U>(s, LIDX(k, i, j - 1))));
  edge = ((b_S + b_N) * k.ry + (b_W + b_E) * k.rx);
  const double gc = G<IS_U>(s, c);
  double b_SW = (i > 0 && j > 0) ? 0.0
              : ((i == 0 && j > 1) ? rbx * (gc - G<IS_U>(s, LIDX(k, i, j - 2)))
              : ((i > 1 && j == 0) ? rby * (gc - G<IS_U>(s, LIDX(k, i - 2, j))) : 0.0));
  double b_SE = (i < (nx - 1) && j > 0) ? 0.0
              : ((i == (nx - 1) && j > 1) ? -rbx * (gc - G<IS_U>(s, LIDX(k, i, j - 2)))
              : ((i < (nx - 2) && j == 0) ? rby * (G<IS_U>(s, LIDX(k, i + 2, j)) - gc) : 0.0));
  double b_NW = (i > 0 && j < (ny - 1)) ? 0.0
              : ((i == 0 && j < (ny - 2)) ? rbx * (G<IS_U>(s, LIDX(k, i, j + 2)) - gc)
              : ((i > 1 && j == (ny - 1)) ? -rby * (gc - G<IS_U>(s, LIDX(k, i - 2, j))) : 0.0));
  double b_NE = (i < (nx - 1) && j < (ny - 1)) ? 0.0
              : ((i == (nx - 1) && j < (ny - 2)) ? -rbx * (G<IS_U>(s, LIDX(k, i, j + 2)) - gc)
              : ((i < (nx - 2) && j == (ny - 1)) ? -rby * (G<IS_U>(s, LIDX(k, i + 2, j)) - gc) : 0.0));
  return (G<IS_U>(s, SW) + b_SW) + (G<IS_U>(s, NE) + b_NE) - (G<IS_U>(s, SE) + b_SE) -
         (G<IS_U>(s, NW) + b_NW);
}

// du2dt / dv2dt of one cell for one stage (reactionDiffusion.cu:147-499).  (i, gj) global.
__device__ void rd_cell(const YhK &k, const Stage &s, const uint8_t *solid, int i, int gj,
                        double u, double v, double I_sum, double I_v, double &du_out,
                        double &dv_out) {
  const int nx = k.nx, ny = k.nyg;
  const int c = LIDX(k, i, gj);
  const double rx = k.rx, ry = k.ry, rs = k.rscale;
  double du = 0.0, dv = 0.0;
  if (k.neumannBC) {
    const int S = LIDX(k, i, yh_mir(gj - 1, ny)), N = LIDX(k, i, yh_mir(gj + 1, ny));
    const int W = LIDX(k, yh_mir(i - 1, nx), gj), E = LIDX(k, yh_mir(i + 1, nx), gj);
    if (k.solidSwitch) {   // :154-184
      const bool sc = solid[c], sw = solid[W], se = solid[E], sn = solid[N], ss = solid[S];
      const float cxx = (sw && se) && (sw && sc) ? 1.0f : ((sw && sc) ? 2.0f : 0.0f);
      const float cxy = sc ? ((sw || se) ? 2.0f : 0.0f) : 0.0f;
      const float cxz = (sw && se) && (sc && se) ? 1.0f : ((sc && se) ? 2.0f : 0.0f);
      const float cyx = (sn && ss) && (sn && sc) ? 1.0f : ((sn && sc) ? 2.0f : 0.0f);
      const float cyy = sc ? ((sn || ss) ? 2.0f : 0.0f) : 0.0f;
      const float cyz = (sn && ss) && (sc && ss) ? 1.0f : ((sc && ss) ? 2.0f : 0.0f);
      du = ((cxx * s.u(W) - cxy * u + cxz * s.u(E)) * rx + (cyx * s.u(N) - cyy * u + cyz * s.u(S)) * ry);
      if (k.gateDiff)
        dv = ((cxx * s.v(W) - cxy * v + cxz * s.v(E)) * rx * rs +
              (cyx * s.v(N) - cyy * v + cyz * s.v(S)) * ry * rs);
    } else {   // :186-356
      const double uW = s.u(W), uE = s.u(E), uN = s.u(N), uS = s.u(S);
      du = ((uW - 2.0 * u + uE) * rx + (uN - 2.0 * u + uS) * ry);
      double vW = 0, vE = 0, vN = 0, vS = 0;
      if (k.gateDiff) {
        vW = s.v(W); vE = s.v(E); vN = s.v(N); vS = s.v(S);
        dv = ((vW - 2.0 * v + vE) * rx * rs + (vN - 2.0 * v + vS) * ry * rs);
      }
      if (k.lap4) {
        int SW, SE, NW, NE;
        diag_idx(k, i, gj, SW, SE, NW, NE);
        const double q = k.qx4 + k.qy4;
        du += -2.0 * q * (+(uW - u + uE) + (uN - u + uS));
        du += q * (s.u(SW) + s.u(SE) + s.u(NW) + s.u(NE));
        // J at a neighbour = I of the neighbour's stage state (:219,:227-229, synchronous)
        const int iW = yh_mir(i - 1, nx), iE = yh_mir(i + 1, nx);
        const int jS = yh_mir(gj - 1, ny), jN = yh_mir(gj + 1, ny);
        double nvW = s.v(W), nvE = s.v(E), nvN = s.v(N), nvS = s.v(S);
        const double JW = yh_Isum(k, uW, nvW, yh_scs(k, iW, gj));
        const double JE = yh_Isum(k, uE, nvE, yh_scs(k, iE, gj));
        const double JN = yh_Isum(k, uN, nvN, yh_scs(k, i, jN));
        const double JS = yh_Isum(k, uS, nvS, yh_scs(k, i, jS));
        du -= ((JW - 2.0 * I_sum + JE) * k.fx4 + (JN - 2.0 * I_sum + JS) * k.fy4);
        if (k.gateDiff) {
          dv += -rs * 2.0 * q * (+(vW - v + vE) + (vN - v + vS));
          dv += rs * q * (s.v(SW) + s.v(SE) + s.v(NW) + s.v(NE));
          const double KW = yh_Iv(k, uW, nvW), KE = yh_Iv(k, uE, nvE);
          const double KN = yh_Iv(k, uN, nvN), KS = yh_Iv(k, uS, nvS);
          dv -= ((KW - 2.0 * I_v + KE) * k.fx4 + (KN - 2.0 * I_v + KS) * k.fy4);
        }
      }
      if (k.anisotropy) {
        int SW, SE, NW, NE;
        diag_idx(k, i, gj, SW, SE, NW, NE);
        double edge;
        double cr = aniso_term<true>(k, s, i, gj, SW, SE, NW, NE, edge);
        du += edge;
        du += (k.rxy * cr);
        if (k.gateDiff) {
          cr = aniso_term<false>(k, s, i, gj, SW, SE, NW, NE, edge);
          dv += edge;
          dv += (k.rxy * cr * rs);
        }
      }
    }
  } else {   // Dirichlet :360-490 (out-of-domain mask = non-tissue; the reference is UB there)
    const double bv = k.boundaryVal;
    bool sc, sw, se, sn, ss;
    if (k.solidSwitch) {
      sc = solid[c];
      sw = solid_at(k, solid, i - 1, gj); se = solid_at(k, solid, i + 1, gj);
      sn = solid_at(k, solid, i, gj + 1); ss = solid_at(k, solid, i, gj - 1);
    } else {
      sc = true; sw = i > 0; se = i < (nx - 1); sn = gj < (ny - 1); ss = gj > 0;
    }
    const double uS = sc && ss ? s.u(LIDX(k, i, gj - 1)) : bv;
    const double uN = sc && sn ? s.u(LIDX(k, i, gj + 1)) : bv;
    const double uW = sc && sw ? s.u(LIDX(k, i - 1, gj)) : bv;
    const double uE = sc && se ? s.u(LIDX(k, i + 1, gj)) : bv;
    du = ((uW - 2.0 * u + uE) * rx + (uN - 2.0 * u + uS) * ry);
    if (k.gateDiff) {
      const double vS = sc && ss ? s.v(LIDX(k, i, gj - 1)) : bv;
      const double vN = sc && sn ? s.v(LIDX(k, i, gj + 1)) : bv;
      const double vW = sc && sw ? s.v(LIDX(k, i - 1, gj)) : bv;
      const double vE = sc && se ? s.v(LIDX(k, i + 1, gj)) : bv;
      dv = ((vW - 2.0 * v + vE) * rx * rs + (vN - 2.0 * v + vS) * ry * rs);
    }
    if (k.anisotropy) {
      bool ssw, sse, snw, sne;
      if (k.solidSwitch) {
        ssw = solid_at(k, solid, i - 1, gj - 1); sse = solid_at(k, solid, i + 1, gj - 1);
        snw = solid_at(k, solid, i - 1, gj + 1); sne = solid_at(k, solid, i + 1, gj + 1);
      } else {
        ssw = (i > 0) && (gj > 0); sse = (i < (nx - 1)) && (gj > 0);
        snw = (i > 0) && (gj < (ny - 1)); sne = (i < (nx - 1)) && (gj < (ny - 1));
      }
      double a = sc && ssw ? s.u(LIDX(k, i - 1, gj - 1)) : bv;
      double b = sc && sse ? s.u(LIDX(k, i + 1, gj - 1)) : bv;
      double cc = sc && snw ? s.u(LIDX(k, i - 1, gj + 1)) : bv;
      double d = sc && sne ? s.u(LIDX(k, i + 1, gj + 1)) : bv;
      du += (k.rxy * (a + d - b - cc));
      if (k.gateDiff) {
        a = sc && ssw ? s.v(LIDX(k, i - 1, gj - 1)) : bv;
        b = sc && sse ? s.v(LIDX(k, i + 1, gj - 1)) : bv;
        cc = sc && snw ? s.v(LIDX(k, i - 1, gj + 1)) : bv;
        d = sc && sne ? s.v(LIDX(k, i + 1, gj + 1)) : bv;
        dv += (k.rxy * (a + d - b - cc) * rs);
      }
    }
  }
  du -= k.dt * I_sum;   // :498-499
  dv -= k.dt * I_v;
  du_out = du;
  dv_out = dv;
}

struct StageArgs {
  const double *u0, *v0;     // state at the start of the step
  const double *Us, *Vs;     // stage state (== u0/v0 with add0 on the first stage)
  double *Un, *Vn;           // next stage state (not LAST)
  double *ru, *rv;           // rhs accumulators (K > 1)
  double *uo, *vo, *vtu, *vtv;   // outputs (LAST)
  const uint8_t *solid;
  double a_next, w;          // ki of the NEXT stage, weight of this one
  int first, last;
  int jlo, jhi;              // local rows computed by this stage
};

__global__ void __launch_bounds__(256)
rd_stage_kernel(const __grid_constant__ YhK k, const __grid_constant__ StageArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = a.jlo + blockIdx.y * blockDim.y + threadIdx.y;   // local row
  if (i >= k.nx || j >= a.jhi) return;
  const int gj = k.jg0 + j;
  const int c = i + k.nx * j;
  Stage s{a.Us, a.Vs, a.first != 0};
  const double u = s.u(c), v = s.v(c);
  const double I_sum = yh_Isum(k, u, v, yh_scs(k, i, gj));
  const double I_v = yh_Iv(k, u, v);
  double du, dv;
  rd_cell(k, s, a.solid, i, gj, u, v, I_sum, I_v, du, dv);
  double ru = 0.0, rv = 0.0;
  if (!a.first) { ru = a.ru[c]; rv = a.rv[c]; }
  ru += (a.w * du);   // :502-503
  rv += (a.w * dv);
  if (!a.last) {
    a.ru[c] = ru; a.rv[c] = rv;
    a.Un[c] = a.u0[c] + (a.a_next * du);   // :117-118 of the next stage
    a.Vn[c] = a.v0[c] + (a.a_next * dv);
  } else {
    double u0 = a.u0[c], v0 = a.v0[c];
    u0 += k.tc * ru;   // :512-513
    v0 += k.tc * rv;
    const bool sc = k.solidSwitch ? a.solid[c] != 0 : true;
    a.uo[c] = sc ? u0 : 0.0;   // :515-561
    a.vo[c] = sc ? v0 : 0.0;
    if (k.gateDiff && a.vtu) {
      a.vtu[c] = sc ? ru / k.dt : 0.0;
      a.vtv[c] = sc ? rv / k.dt : 0.0;
    }
  }
}

}  // namespace

int yh_launch_rd_generic(const YhK &k, const double *u_in, const double *v_in, double *u_out,
                         double *v_out, double *vtu, double *vtv, const uint8_t *solid,
                         cudaStream_t st) {
  const int K = k.timeIntOrder;
  double ki[4] = {0, 0, 0, 0}, w[4] = {0, 0, 0, 0};
  switch (K) {   // reactionDiffusion.cu:71-93
    case 1: w[0] = 1.0; break;
    case 2: ki[1] = 0.5; w[1] = 1.0; break;
    case 4: ki[1] = 0.5; ki[2] = 0.5; ki[3] = 1.0;
            w[0] = 0.166666666666667; w[1] = 0.333333333333333;
            w[2] = 0.333333333333333; w[3] = 0.166666666666667; break;
    default: yh_set_error("timeIntOrder must be 1, 2 or 4"); return YH_ERR_INVALID_ARG;
  }
  const size_t n = (size_t)k.nx * k.ny;
  double *ws = nullptr;
  if (K > 1) {
    int rc = yh_workspace(6 * n * sizeof(double), (void **)&ws, 0);
    if (rc != YH_OK) return rc;
  }
  double *S[2][2] = {{ws, ws ? ws + n : nullptr}, {ws ? ws + 2 * n : nullptr, ws ? ws + 3 * n : nullptr}};
  double *ru = ws ? ws + 4 * n : nullptr, *rv = ws ? ws + 5 * n : nullptr;
  // rows of the local array that exist in the global domain
  const int lo = 0, hi = k.ny;
  for (int s = 0; s < K; s++) {
    StageArgs a;
    a.u0 = u_in; a.v0 = v_in;
    a.first = (s == 0); a.last = (s == K - 1);
    a.Us = a.first ? u_in : S[(s - 1) & 1][0];
    a.Vs = a.first ? v_in : S[(s - 1) & 1][1];
    a.Un = a.last ? nullptr : S[s & 1][0];
    a.Vn = a.last ? nullptr : S[s & 1][1];
    a.ru = ru; a.rv = rv;
    a.uo = u_out; a.vo = v_out; a.vtu = vtu; a.vtv = vtv;
    a.solid = solid;
    a.a_next = a.last ? 0.0 : ki[s + 1];
    a.w = w[s];
    // later stages read `rad` more rows of this stage's output per remaining stage (rad = 2 for the
    // anisotropic no-flux corner terms, reactionDiffusion.cu:290-304)
    const int rad = (k.anisotropy && k.neumannBC && !k.solidSwitch) ? 2 : 1;
    const int ext = rad * (K - 1 - s);
    a.jlo = max(lo, k.row0 - ext);
    a.jhi = min(hi, k.row1 + ext);
    if (a.jhi <= a.jlo) continue;
    dim3 blk(32, 8), grd((k.nx + 31) / 32, (a.jhi - a.jlo + 7) / 8);
    YH_LAUNCH(rd_stage_kernel, grd, blk, 0, st, k, a);
  }
  return YH_OK;
}
