/*
 * yh_oracle.c -- TEST INFRASTRUCTURE (see yh_oracle.h).  Plain-C restatement of the
 * Yolohtli monodomain hot path; every function cites the reference lines it follows.
 * Expression order is kept literally so that, with contraction off on both sides
 * (gcc -ffp-contract=off / nvcc --fmad=false), results are bit-identical to the
 * reference's kernels in the race-free modes.
 */
#include "yh_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define I2D(nn, i, j) (((nn) * (j)) + (i)) /* typeDefinition.cuh: I2D */

static int g_threads = 0;
void yho_set_threads(int n) {
  g_threads = n;
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}
int yho_get_threads(void) {
#ifdef _OPENMP
  return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
  return 1;
#endif
}

/* helper_functions.cu:69-79 coord_i / coord_j: Neumann mirror index */
static inline int mir(int i, int n) {
  return (int)((i >= 0) && (i < n)) * i + (int)(i < 0) * (-i) + (int)(i >= n) * (2 * (n - 1) - i);
}

/* ------------------------------------------------------------------------------------
 * Parameters: parameterSetup(), saveFiles.cu:105-231, then main.cu:148-158.
 * ---------------------------------------------------------------------------------- */
int yho_params_derive(yh_params *p, double Dxx, double Dyy, double Dxy) {
  /* saveFiles.cu:153-170 / main.cu:149-155 */
  p->rxy = 2.0 * Dxy * p->dt / (4.0 * p->hx * p->hy);
  p->rbx = p->hx * Dxy / (Dxx * p->hy);
  p->rby = p->hy * Dxy / (Dyy * p->hx);
  p->rx = p->dt * Dxx / (p->hx * p->hx);
  p->ry = p->dt * Dyy / (p->hy * p->hy);
  p->invdx = 0.5 / p->hx;
  p->invdy = 0.5 / p->hy;
  p->qx4 = p->dt * Dyy / (p->hy * p->hy * 12.0);
  p->qy4 = p->dt * Dxx / (p->hx * p->hx * 12.0);
  p->fx4 = p->dt / 12.0;
  p->fy4 = p->dt / 12.0;
  return 0;
}

int yho_params_default(yh_params *p, int nx, int ny, int reduce_sym, int scale_L) {
  if (!p || nx < 4 || ny < 4) return YH_ERR_INVALID_ARG;
  memset(p, 0, sizeof(*p));
  p->nx = nx; p->ny = ny; p->ny_global = ny; p->jg0 = 0;
  p->solidSwitch = 0; p->neumannBC = 1; p->gateDiff = 1; p->tipAlgorithm = 1;
  p->anisotropy = 0; p->tipGrad = 0; p->lap4 = 4; p->timeIntOrder = 4;
  /* saveFiles.cu:140-146; scale_L keeps hx at the 512-grid value */
  p->Lx = scale_L ? 12.0 * (nx - 1.0) / 511.0 : 12.0;
  p->Ly = scale_L ? 12.0 * (ny - 1.0) / 511.0 : 12.0;
  p->hx = p->Lx / (nx - 1.0);
  p->hy = p->Ly / (ny - 1.0);
  p->dt = 0.02;
  double diff_par = 0.001, diff_per = 0.001, degrad = 0.0;
  double th = degrad * 3.14159265359 / 180.0; /* typeDefinition.cuh: pi */
  double Dxx = diff_par * cos(th) * cos(th) + diff_per * sin(th) * sin(th);
  double Dyy = diff_par * sin(th) * sin(th) + diff_per * cos(th) * cos(th);
  double Dxy = (diff_par - diff_per) * sin(th) * cos(th);
  p->rscale = 0.01;
  p->tipOffsetX = 160; p->tipOffsetY = 160;
  p->tipx0 = 0.0f; p->tipy0 = 0.0f;
  p->boundaryVal = 0.0;
  p->Uth = 0.7;
  p->tc = 1.0; p->alpha = 0.2; p->beta = 1.1; p->gamma = 0.0; p->delta = 1.0;
  p->eps = 0.005; p->mu = 1.0; p->theta = 0.0;
  /* main.cu:148: dt halved in symmetry-reduction mode, then the r's recomputed */
  p->dt = reduce_sym ? 0.5 * p->dt : p->dt;
  return yho_params_derive(p, Dxx, Dyy, Dxy);
}

/* main.cu:606-618 */
void yho_cross_field_ic(int nx, int ny, double *u, double *v) {
  memset(u, 0, sizeof(double) * (size_t)nx * ny);
  memset(v, 0, sizeof(double) * (size_t)nx * ny);
  for (int j = 0; j < ny; j++)
    for (int i = 0; i < nx / 8; i++) u[i + (size_t)nx * j] = 1.0;
  for (int j = ny / 2; j < ny; j++)
    for (int i = 0; i < nx; i++) v[i + (size_t)nx * j] = 1.0;
}

/* ------------------------------------------------------------------------------------
 * Reaction-diffusion step, reactionDiffusion.cu:26-566, synchronous stages.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  const yh_params *p;
  const double *U, *V;     /* stage state */
  const double *Ju, *Jv;   /* pointwise I_sum / I_v of the stage state (lap4 only) */
  const uint8_t *solid;
} rd_ctx;

/* The four diagonal indices with per-axis mirror, reactionDiffusion.cu:203-217 */
static inline void diag_idx(int nx, int ny, int i, int j, int *SW, int *SE, int *NW, int *NE) {
  *SW = (i > 0 && j > 0) ? I2D(nx, i - 1, j - 1)
      : ((i == 0 && j > 0) ? I2D(nx, i + 1, j - 1)
      : ((i > 0 && j == 0) ? I2D(nx, i - 1, j + 1) : I2D(nx, i + 1, j + 1)));
  *SE = (i < (nx - 1) && j > 0) ? I2D(nx, i + 1, j - 1)
      : ((i == (nx - 1) && j > 0) ? I2D(nx, i - 1, j - 1)
      : ((i < (nx - 1) && j == 0) ? I2D(nx, i + 1, j + 1) : I2D(nx, i - 1, j + 1)));
  *NW = (i > 0 && j < (ny - 1)) ? I2D(nx, i - 1, j + 1)
      : ((i == 0 && j < (ny - 1)) ? I2D(nx, i + 1, j + 1)
      : ((i > 0 && j == (ny - 1)) ? I2D(nx, i - 1, j - 1) : I2D(nx, i + 1, j - 1)));
  *NE = (i < (nx - 1) && j < (ny - 1)) ? I2D(nx, i + 1, j + 1)
      : ((i == (nx - 1) && j < (ny - 1)) ? I2D(nx, i - 1, j + 1)
      : ((i < (nx - 1) && j == (ny - 1)) ? I2D(nx, i + 1, j - 1) : I2D(nx, i - 1, j - 1)));
}

/* anisotropy edge/corner corrections for one field, reactionDiffusion.cu:267-307 */
static inline double aniso_term(const yh_params *p, const double *g, int i, int j,
                                int SW, int SE, int NW, int NE, double *edge) {
  const int nx = p->nx, ny = p->ny;
  const int i2d = I2D(nx, i, j);
  const double rbx = p->rbx, rby = p->rby;
  double b_S = (j > 0) ? 0.0
             : ((j == 0 && (i == 0 || i == (nx - 1))) ? 0.0
             : rby * (g[I2D(nx, i + 1, j)] - g[I2D(nx, i - 1, j)]));
  double b_N = (j < (ny - 1)) ? 0.0
             : ((j == (ny - 1) && (i == 0 || i == (nx - 1))) ? 0.0
             : -rby * (g[I2D(nx, i + 1, j)] - g[I2D(nx, i - 1, j)]));
  double b_W = (i > 0) ? 0.0
             : ((i == 0 && (j == 0 || j == (ny - 1))) ? 0.0
             : rbx * (g[I2D(nx, i, j + 1)] - g[I2D(nx, i, j - 1)]));
  double b_E = (i < (nx - 1)) ? 0.0
             : ((i == (nx - 1) && (j == 0 || j == (ny - 1))) ? 0.0
             : -rbx * (g[I2D(nx, i, j + 1)] - g[I2D(nx, i, j - 1)]));
  *edge = ((b_S + b_N) * p->ry + (b_W + b_E) * p->rx);

  double b_SW = (i > 0 && j > 0) ? 0.0
              : ((i == 0 && j > 1) ? rbx * (g[i2d] - g[I2D(nx, i, j - 2)])
              : ((i > 1 && j == 0) ? rby * (g[i2d] - g[I2D(nx, i - 2, j)]) : 0.0));
  double b_SE = (i < (nx - 1) && j > 0) ? 0.0
              : ((i == (nx - 1) && j > 1) ? -rbx * (g[i2d] - g[I2D(nx, i, j - 2)])
              : ((i < (nx - 2) && j == 0) ? rby * (g[I2D(nx, i + 2, j)] - g[i2d]) : 0.0));
  double b_NW = (i > 0 && j < (ny - 1)) ? 0.0
              : ((i == 0 && j < (ny - 2)) ? rbx * (g[I2D(nx, i, j + 2)] - g[i2d])
              : ((i > 1 && j == (ny - 1)) ? -rby * (g[i2d] - g[I2D(nx, i - 2, j)]) : 0.0));
  double b_NE = (i < (nx - 1) && j < (ny - 1)) ? 0.0
              : ((i == (nx - 1) && j < (ny - 2)) ? -rbx * (g[I2D(nx, i, j + 2)] - g[i2d])
              : ((i < (nx - 2) && j == (ny - 1)) ? -rby * (g[I2D(nx, i + 2, j)] - g[i2d]) : 0.0));
  return (g[SW] + b_SW) + (g[NE] + b_NE) - (g[SE] + b_SE) - (g[NW] + b_NW);
}

/* mask lookup with out-of-domain cells counted as non-tissue (the reference indexes
 * unclamped in the Dirichlet branches, reactionDiffusion.cu:362-373: UB at the edges). */
static inline int solid_at(const uint8_t *solid, int nx, int ny, int i, int j) {
  return (i >= 0 && i < nx && j >= 0 && j < ny) ? (solid[I2D(nx, i, j)] != 0) : 0;
}

static void rd_cell(const rd_ctx *c, int i, int j, double I_sum, double I_v,
                    double *du_out, double *dv_out) {
  const yh_params *p = c->p;
  const int nx = p->nx, ny = p->ny;
  const double *gu = c->U, *gv = c->V;
  const int i2d = I2D(nx, i, j);
  const double u = gu[i2d], v = gv[i2d];
  const double rx_d = p->rx, ry_d = p->ry, rscale_d = p->rscale;
  double du2dt = 0.0, dv2dt = 0.0;

  if (p->neumannBC) {
    int S = I2D(nx, i, mir(j - 1, ny));   /* :149-152 */
    int N = I2D(nx, i, mir(j + 1, ny));
    int W = I2D(nx, mir(i - 1, nx), j);
    int E = I2D(nx, mir(i + 1, nx), j);
    if (p->solidSwitch) {
      int sc = c->solid[i2d] != 0, sw = c->solid[W] != 0, se = c->solid[E] != 0;
      int sn = c->solid[N] != 0, ss = c->solid[S] != 0;
      /* :162-169, float3 holding exact 0/1/2 */
      float cxx = (sw && se) && (sw && sc) ? 1.0f : ((sw && sc) ? 2.0f : 0.0f);
      float cxy = sc ? ((sw || se) ? 2.0f : 0.0f) : 0.0f;
      float cxz = (sw && se) && (sc && se) ? 1.0f : ((sc && se) ? 2.0f : 0.0f);
      float cyx = (sn && ss) && (sn && sc) ? 1.0f : ((sn && sc) ? 2.0f : 0.0f);
      float cyy = sc ? ((sn || ss) ? 2.0f : 0.0f) : 0.0f;
      float cyz = (sn && ss) && (sc && ss) ? 1.0f : ((sc && ss) ? 2.0f : 0.0f);
      du2dt = ((cxx * gu[W] - cxy * u + cxz * gu[E]) * rx_d
             + (cyx * gu[N] - cyy * u + cyz * gu[S]) * ry_d);   /* :171-173 */
      if (p->gateDiff) {
        dv2dt = ((cxx * gv[W] - cxy * v + cxz * gv[E]) * rx_d * rscale_d
               + (cyx * gv[N] - cyy * v + cyz * gv[S]) * ry_d * rscale_d);   /* :178-180 */
      }
    } else {
      du2dt = ((gu[W] - 2.0 * u + gu[E]) * rx_d + (gu[N] - 2.0 * u + gu[S]) * ry_d);  /* :188-190 */
      if (p->gateDiff) {
        dv2dt = ((gv[W] - 2.0 * v + gv[E]) * rx_d * rscale_d
               + (gv[N] - 2.0 * v + gv[S]) * ry_d * rscale_d);   /* :195-197 */
      }
      if (p->lap4) {   /* :201-247 */
        int SW, SE, NW, NE;
        diag_idx(nx, ny, i, j, &SW, &SE, &NW, &NE);
        const double qx4 = p->qx4, qy4 = p->qy4;
        du2dt += -2.0 * (qx4 + qy4) * (+(gu[W] - u + gu[E]) + (gu[N] - u + gu[S]));
        du2dt += (qx4 + qy4) * (gu[SW] + gu[SE] + gu[NW] + gu[NE]);
        du2dt -= ((c->Ju[W] - 2.0 * I_sum + c->Ju[E]) * p->fx4
                + (c->Ju[N] - 2.0 * I_sum + c->Ju[S]) * p->fy4);
        if (p->gateDiff) {
          dv2dt += -rscale_d * 2.0 * (qx4 + qy4) * (+(gv[W] - v + gv[E]) + (gv[N] - v + gv[S]));
          dv2dt += rscale_d * (qx4 + qy4) * (gv[SW] + gv[SE] + gv[NW] + gv[NE]);
          dv2dt -= ((c->Jv[W] - 2.0 * I_v + c->Jv[E]) * p->fx4
                  + (c->Jv[N] - 2.0 * I_v + c->Jv[S]) * p->fy4);
        }
      }
      if (p->anisotropy) {   /* :249-356 */
        int SW, SE, NW, NE;
        diag_idx(nx, ny, i, j, &SW, &SE, &NW, &NE);
        double edge;
        double cr = aniso_term(p, gu, i, j, SW, SE, NW, NE, &edge);
        du2dt += edge;
        du2dt += (p->rxy * cr);
        if (p->gateDiff) {
          cr = aniso_term(p, gv, i, j, SW, SE, NW, NE, &edge);
          dv2dt += edge;
          dv2dt += (p->rxy * cr * rscale_d);
        }
      }
    }
  } else {   /* Dirichlet, :360-490 */
    const double bv = p->boundaryVal;
    if (p->solidSwitch) {
      int sc = c->solid[i2d] != 0;
      int sw = solid_at(c->solid, nx, ny, i - 1, j), se = solid_at(c->solid, nx, ny, i + 1, j);
      int sn = solid_at(c->solid, nx, ny, i, j + 1), ss = solid_at(c->solid, nx, ny, i, j - 1);
      double uS = sc && ss ? gu[I2D(nx, i, j - 1)] : bv;
      double uN = sc && sn ? gu[I2D(nx, i, j + 1)] : bv;
      double uW = sc && sw ? gu[I2D(nx, i - 1, j)] : bv;
      double uE = sc && se ? gu[I2D(nx, i + 1, j)] : bv;
      du2dt = ((uW - 2.0 * u + uE) * rx_d + (uN - 2.0 * u + uS) * ry_d);
      if (p->gateDiff) {
        double vS = sc && ss ? gv[I2D(nx, i, j - 1)] : bv;
        double vN = sc && sn ? gv[I2D(nx, i, j + 1)] : bv;
        double vW = sc && sw ? gv[I2D(nx, i - 1, j)] : bv;
        double vE = sc && se ? gv[I2D(nx, i + 1, j)] : bv;
        dv2dt = ((vW - 2.0 * v + vE) * rx_d * rscale_d + (vN - 2.0 * v + vS) * ry_d * rscale_d);
      }
      if (p->anisotropy) {
        int ssw = solid_at(c->solid, nx, ny, i - 1, j - 1), sse = solid_at(c->solid, nx, ny, i + 1, j - 1);
        int snw = solid_at(c->solid, nx, ny, i - 1, j + 1), sne = solid_at(c->solid, nx, ny, i + 1, j + 1);
        double a = sc && ssw ? gu[I2D(nx, i - 1, j - 1)] : bv;
        double b = sc && sse ? gu[I2D(nx, i + 1, j - 1)] : bv;
        double cc = sc && snw ? gu[I2D(nx, i - 1, j + 1)] : bv;
        double d = sc && sne ? gu[I2D(nx, i + 1, j + 1)] : bv;
        du2dt += (p->rxy * (a + d - b - cc));
        if (p->gateDiff) {
          a = sc && ssw ? gv[I2D(nx, i - 1, j - 1)] : bv;
          b = sc && sse ? gv[I2D(nx, i + 1, j - 1)] : bv;
          cc = sc && snw ? gv[I2D(nx, i - 1, j + 1)] : bv;
          d = sc && sne ? gv[I2D(nx, i + 1, j + 1)] : bv;
          dv2dt += (p->rxy * (a + d - b - cc) * rscale_d);
        }
      }
    } else {
      double uS = j > 0 ? gu[I2D(nx, i, j - 1)] : bv;
      double uN = j < (ny - 1) ? gu[I2D(nx, i, j + 1)] : bv;
      double uW = i > 0 ? gu[I2D(nx, i - 1, j)] : bv;
      double uE = i < (nx - 1) ? gu[I2D(nx, i + 1, j)] : bv;
      du2dt = ((uW - 2.0 * u + uE) * rx_d + (uN - 2.0 * u + uS) * ry_d);
      if (p->gateDiff) {
        double vS = j > 0 ? gv[I2D(nx, i, j - 1)] : bv;
        double vN = j < (ny - 1) ? gv[I2D(nx, i, j + 1)] : bv;
        double vW = i > 0 ? gv[I2D(nx, i - 1, j)] : bv;
        double vE = i < (nx - 1) ? gv[I2D(nx, i + 1, j)] : bv;
        dv2dt = ((vW - 2.0 * v + vE) * rx_d * rscale_d + (vN - 2.0 * v + vS) * ry_d * rscale_d);
      }
      if (p->anisotropy) {
        double a = (i > 0) && (j > 0) ? gu[I2D(nx, i - 1, j - 1)] : bv;
        double b = (i < (nx - 1)) && (j > 0) ? gu[I2D(nx, i + 1, j - 1)] : bv;
        double cc = (i > 0) && (j < (ny - 1)) ? gu[I2D(nx, i - 1, j + 1)] : bv;
        double d = (i < (nx - 1)) && (j < (ny - 1)) ? gu[I2D(nx, i + 1, j + 1)] : bv;
        du2dt += (p->rxy * (a + d - b - cc));
        if (p->gateDiff) {
          a = (i > 0) && (j > 0) ? gv[I2D(nx, i - 1, j - 1)] : bv;
          b = (i < (nx - 1)) && (j > 0) ? gv[I2D(nx, i + 1, j - 1)] : bv;
          cc = (i > 0) && (j < (ny - 1)) ? gv[I2D(nx, i - 1, j + 1)] : bv;
          d = (i < (nx - 1)) && (j < (ny - 1)) ? gv[I2D(nx, i + 1, j + 1)] : bv;
          dv2dt += (p->rxy * (a + d - b - cc) * rscale_d);
        }
      }
    }
  }
  du2dt -= p->dt * I_sum;   /* :498-499 */
  dv2dt -= p->dt * I_v;
  *du_out = du2dt;
  *dv_out = dv2dt;
}

int yho_rd_step(const yh_params *p, const double *u_in, const double *v_in,
                double *u_out, double *v_out, double *velTan_u, double *velTan_v,
                const uint8_t *solid, int stim_mouse, int point_x, int point_y) {
  if (!p || !u_in || !v_in || !u_out || !v_out) return YH_ERR_INVALID_ARG;
  if (p->solidSwitch && !solid) return YH_ERR_INVALID_ARG;
  if (p->jg0 != 0 || p->ny_global != p->ny) return YH_ERR_UNSUPPORTED;
  const int nx = p->nx, ny = p->ny;
  const size_t n = (size_t)nx * ny;
  double ki[4] = {0, 0, 0, 0}, w[4] = {0, 0, 0, 0};
  int K = p->timeIntOrder;
  switch (K) {   /* :71-93 */
    case 1: w[0] = 1.0; break;
    case 2: ki[1] = 0.5; w[1] = 1.0; break;
    case 4: ki[1] = 0.5; ki[2] = 0.5; ki[3] = 1.0;
            w[0] = 0.166666666666667; w[1] = 0.333333333333333;
            w[2] = 0.333333333333333; w[3] = 0.166666666666667; break;
    default: return YH_ERR_INVALID_ARG;
  }
  const int needJ = p->neumannBC && !p->solidSwitch && p->lap4;
  double *U = (double *)malloc(n * sizeof(double)), *V = (double *)malloc(n * sizeof(double));
  double *du = (double *)calloc(n, sizeof(double)), *dv = (double *)calloc(n, sizeof(double));
  double *ru = (double *)calloc(n, sizeof(double)), *rv = (double *)calloc(n, sizeof(double));
  double *Ju = (double *)malloc(n * sizeof(double)), *Jv = (double *)malloc(n * sizeof(double));
  if (!U || !V || !du || !dv || !ru || !rv || !Ju || !Jv) return YH_ERR_INVALID_ARG;

  for (int k = 0; k < K; k++) {
    /* :117-121 stage state; :131-141 reaction terms */
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; j++) {
      for (int i = 0; i < nx; i++) {
        size_t c = (size_t)i + (size_t)nx * j;
        double u = u_in[c] + (ki[k] * du[c]);
        double v = v_in[c] + (ki[k] * dv[c]);
        U[c] = u; V[c] = v;
        int ic = i - nx / 2, jc = j - ny / 2;            /* :56-61 */
        int cx = point_x - nx / 2, cy = point_y - ny / 2;
        int scs = (((ic - cx) * (ic - cx) + (jc - cy) * (jc - cy)) < 400) && stim_mouse;
        Ju[c] = -(p->mu * u * (1.0 - u) * (u - p->alpha) - u * v) - (scs ? 24.7 : 0.0);
        Jv[c] = -(p->eps * (p->delta * (u - p->gamma) * (p->beta - u) - v - p->theta));
      }
    }
    rd_ctx ctx = {p, U, V, needJ ? Ju : NULL, needJ ? Jv : NULL, solid};
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; j++) {
      for (int i = 0; i < nx; i++) {
        size_t c = (size_t)i + (size_t)nx * j;
        double a, b;
        rd_cell(&ctx, i, j, Ju[c], Jv[c], &a, &b);
        du[c] = a; dv[c] = b;
        ru[c] += (w[k] * a);   /* :502-503 */
        rv[c] += (w[k] * b);
      }
    }
  }
#pragma omp parallel for schedule(static)
  for (int j = 0; j < ny; j++) {
    for (int i = 0; i < nx; i++) {
      size_t c = (size_t)i + (size_t)nx * j;
      double u0 = u_in[c], v0 = v_in[c];
      u0 += p->tc * ru[c];   /* :512-513 */
      v0 += p->tc * rv[c];
      int sc = p->solidSwitch ? (solid[c] != 0) : 1;
      u_out[c] = sc ? u0 : 0.0;   /* :515-561 */
      v_out[c] = sc ? v0 : 0.0;
      if (p->gateDiff && velTan_u && velTan_v) {
        velTan_u[c] = sc ? ru[c] / p->dt : 0.0;
        velTan_v[c] = sc ? rv[c] / p->dt : 0.0;
      }
    }
  }
  free(U); free(V); free(du); free(dv); free(ru); free(rv); free(Ju); free(Jv);
  return YH_OK;
}

/* Fast path for the headline mode (Euler, 5-point, Neumann, square domain): same expression
 * order as rd_cell()/yho_rd_step(), fused into one sweep so the cpu_baseline timing is a fair
 * "plain-C host loop".  Checked against yho_rd_step() in tests/test_oracle.py. */
static void rd_euler5_sweep(const yh_params *p, const double *u_in, const double *v_in,
                            double *u_out, double *v_out, int stim_mouse, int px, int py) {
  const int nx = p->nx, ny = p->ny;
  const double rx = p->rx, ry = p->ry, rs = p->rscale, dt = p->dt, tc = p->tc;
  const int gd = p->gateDiff;
#pragma omp parallel for schedule(static)
  for (int j = 0; j < ny; j++) {
    const double *uc = u_in + (size_t)nx * j, *vc = v_in + (size_t)nx * j;
    const double *un = u_in + (size_t)nx * mir(j + 1, ny), *us = u_in + (size_t)nx * mir(j - 1, ny);
    const double *vn = v_in + (size_t)nx * mir(j + 1, ny), *vs = v_in + (size_t)nx * mir(j - 1, ny);
    double *uo = u_out + (size_t)nx * j, *vo = v_out + (size_t)nx * j;
    for (int i = 0; i < nx; i++) {
      int iw = mir(i - 1, nx), ie = mir(i + 1, nx);
      double u0 = uc[i], v0 = vc[i];
      double u = u0 + (0.0 * 0.0), v = v0 + (0.0 * 0.0);
      int scs = 0;
      if (stim_mouse) {
        int ic = i - nx / 2, jc = j - ny / 2, cx = px - nx / 2, cy = py - ny / 2;
        scs = ((ic - cx) * (ic - cx) + (jc - cy) * (jc - cy)) < 400;
      }
      double I_sum = -(p->mu * u * (1.0 - u) * (u - p->alpha) - u * v) - (scs ? 24.7 : 0.0);
      double I_v = -(p->eps * (p->delta * (u - p->gamma) * (p->beta - u) - v - p->theta));
      double du = ((uc[iw] - 2.0 * u + uc[ie]) * rx + (un[i] - 2.0 * u + us[i]) * ry);
      double dv = 0.0;
      if (gd) dv = ((vc[iw] - 2.0 * v + vc[ie]) * rx * rs + (vn[i] - 2.0 * v + vs[i]) * ry * rs);
      du -= dt * I_sum;
      dv -= dt * I_v;
      double ru = 0.0, rv = 0.0;
      ru += (1.0 * du);
      rv += (1.0 * dv);
      uo[i] = u0 + tc * ru;
      vo[i] = v0 + tc * rv;
    }
  }
}

int yho_rd_advance(const yh_params *p, int nsteps, double *u, double *v,
                   const uint8_t *solid, int stim_mouse, int point_x, int point_y) {
  if (!p || !u || !v || nsteps < 0) return YH_ERR_INVALID_ARG;
  const size_t n = (size_t)p->nx * p->ny;
  double *u2 = (double *)malloc(n * sizeof(double)), *v2 = (double *)malloc(n * sizeof(double));
  if (!u2 || !v2) return YH_ERR_INVALID_ARG;
  double *a_u = u, *a_v = v, *b_u = u2, *b_v = v2;
  const int fast = p->timeIntOrder == 1 && !p->lap4 && p->neumannBC && !p->solidSwitch &&
                   !p->anisotropy && p->jg0 == 0 && p->ny_global == p->ny;
  int rc = YH_OK;
  for (int s = 0; s < nsteps && rc == YH_OK; s++) {
    if (fast) rd_euler5_sweep(p, a_u, a_v, b_u, b_v, stim_mouse, point_x, point_y);
    else rc = yho_rd_step(p, a_u, a_v, b_u, b_v, NULL, NULL, solid, stim_mouse, point_x, point_y);
    double *t = a_u; a_u = b_u; b_u = t;   /* swapSoA, helper_functions.cu:140 */
    t = a_v; a_v = b_v; b_v = t;
  }
  if (a_u != u) { memcpy(u, a_u, n * sizeof(double)); memcpy(v, a_v, n * sizeof(double)); }
  free(u2); free(v2);
  return rc;
}

/* ------------------------------------------------------------------------------------
 * Tip tracking, tipTracker.cu:20-517
 * ---------------------------------------------------------------------------------- */
static int equals_tol(double a, double b, double tol) {   /* helper_functions.cu:58-62 */
  return (a == b) || ((a <= (b + tol)) && (a >= (b - tol)));
}

/* gradient(), tipTracker.cu:519-566, indices kept literally */
static void tip_gradient(const yh_params *p, int i, int j, double s, double t,
                         const double *g, float *gx, float *gy) {
  const int nx = p->nx, ny = p->ny;
  int S = (j > 0) ? I2D(nx, i, j - 1) : I2D(nx, i, j + 1);
  int Sx = ((j > 0) && (i < (nx - 1))) ? I2D(nx, i + 1, j - 1) : I2D(nx, i - 1, j + 1);
  int Sy = I2D(nx, i, j);
  int Sxy = (i < (nx - 1)) ? I2D(nx, i + 1, j) : I2D(nx, i - 1, j);
  int N = (j < (ny - 1)) ? I2D(nx, i, j + 1) : I2D(nx, i, j - 1);
  int Nx = ((i < (nx - 1)) && (j < (ny - 1))) ? I2D(nx, i + 1, j + 1) : I2D(nx, i - 1, j - 1);
  int Ny = (j < (ny - 2)) ? I2D(nx, i, j + 2) : ((j == (ny - 2)) ? I2D(nx, i, j) : I2D(nx, i, j - 1));
  int Nxy = ((i < (nx - 1)) && (j < (ny - 2))) ? I2D(nx, i + 1, j + 2)
          : ((j == (ny - 2)) ? I2D(nx, i - 1, j) : I2D(nx, i - 1, j - 1));
  int W = (i > 0) ? I2D(nx, i - 1, j) : I2D(nx, i - 1, j);
  int Wx = I2D(nx, i, j);
  int Wy = ((i > 0) && (j < (ny - 1))) ? I2D(nx, i - 1, j + 1) : I2D(nx, i + 1, j - 1);
  int Wxy = (j < (ny - 1)) ? I2D(nx, i, j + 1) : I2D(nx, i - 1, j);
  int E = (i < (nx - 1)) ? I2D(nx, i + 1, j) : I2D(nx, i - 1, j);
  int Ex = (i < (nx - 2)) ? I2D(nx, i + 2, j) : ((i == (nx - 2)) ? I2D(nx, i, j) : I2D(nx, i - 1, j));
  int Ey = ((i < (nx - 1)) && (j < (ny - 1))) ? I2D(nx, i + 1, j + 1) : I2D(nx, i - 1, j - 1);
  int Exy = ((i < (nx - 2)) && (j < (ny - 1))) ? I2D(nx, i + 2, j + 1)
          : ((i == (nx - 2)) ? I2D(nx, i, j - 1) : I2D(nx, i - 1, j - 1));
  double gx1 = (g[E] - g[W]) * p->invdx, gy1 = (g[N] - g[S]) * p->invdy;
  double gx2 = (g[Ex] - g[Wx]) * p->invdx, gy2 = (g[Nx] - g[Sx]) * p->invdy;
  double gx3 = (g[Ey] - g[Wy]) * p->invdx, gy3 = (g[Ny] - g[Sy]) * p->invdy;
  double gx4 = (g[Exy] - g[Wxy]) * p->invdx, gy4 = (g[Nxy] - g[Sxy]) * p->invdy;
  *gx = (float)((1.0 - s) * (1.0 - t) * gx1 + s * (1.0 - t) * gx2 + t * (1.0 - s) * gx3 + s * t * gx4);
  *gy = (float)((1.0 - s) * (1.0 - t) * gy1 + s * (1.0 - t) * gy2 + t * (1.0 - s) * gy3 + s * t * gy4);
}

/* tipRecordPlane, tipTracker.cu:210-241.  Returns 1 when a record was appended. */
static int tip_record(const yh_params *p, int i, int j, float tx, float ty, double pTime,
                      const double *g_p, uint8_t *tip_plot, int *count, yh_tip *vec, int cap) {
  if (((tx > 0.0) && (tx < 1.0)) && ((ty > 0.0) && (ty < 1.0))) {
    float gx = 0.0f, gy = 0.0f;
    if (p->tipGrad) tip_gradient(p, i, j, tx, ty, g_p, &gx, &gy);
    yh_tip d;
    d.x = (float)(i + tx); d.y = (float)(j + ty); d.vx = gx; d.vy = gy; d.t = (float)pTime;
    if (*count < cap) vec[*count] = d;
    (*count)++;
    if (tip_plot) {   /* plot_field, helper_functions.cu:45-51 */
      int xi = (int)floor(d.x), yi = (int)floor(d.y);
      tip_plot[I2D(p->nx, xi, yi)] = 1;
    }
    return 1;
  }
  return 0;
}

int yho_tip_track(const yh_params *p, const double *g_past, const double *g_present,
                  uint8_t *tip_plot, int *tip_count, yh_tip *tip_vector, int capacity,
                  double pTime, int algorithm) {
  if (!p || !g_past || !g_present || !tip_count || !tip_vector) return YH_ERR_INVALID_ARG;
  const int nx = p->nx, ny = p->ny;
  const double Uth = p->Uth;
  *tip_count = 0;   /* tip_wrapper: cudaMemset(tip_count,0), tipTracker.cu:573 */
  for (int j = 0; j < ny; j++) {
    for (int i = 0; i < nx; i++) {
      int s0 = I2D(nx, i, j);
      int sx = (i < (nx - 1)) ? I2D(nx, i + 1, j) : I2D(nx, i, j);
      int sy = (j < (ny - 1)) ? I2D(nx, i, j + 1) : I2D(nx, i, j);
      int sxy = ((j < (ny - 1)) && (i < (nx - 1))) ? I2D(nx, i + 1, j + 1) : I2D(nx, i, j);
      if (algorithm == 3) {   /* abouzarTip_kernel + abubuFilament, :434-517 */
        double v0 = g_present[s0], vx = g_present[sx], vy = g_present[sy], vxy = g_present[sxy];
        double f0 = v0 - Uth, fx = vx - Uth, fy = vy - Uth, fxy = vxy - Uth;
        double s = (0.f >= f0) + (0.f >= fx) + (0.f >= fy) + (0.f >= fxy);   /* STEP(a,b)=(a>=b) */
        int bv = (s > 0.5f) && (s < 3.5f);
        double d0 = v0 - g_past[s0], dx = vx - g_past[sx], dy = vy - g_past[sy], dxy = vxy - g_past[sxy];
        s = (0.f >= d0) + (0.f >= dx) + (0.f >= dy) + (0.f >= dxy);
        int bdv = (s > 0.5f) && (s < 3.5f);
        if (tip_plot) tip_plot[s0] = (uint8_t)((tip_plot[s0] + (bdv && bv)) != 0);
        continue;
      }
      int inside;
      if (p->solidSwitch) {
        int ic = i - nx / 2, jc = j - ny / 2;
        inside = (ic * ic + jc * jc) < p->tipOffsetX * p->tipOffsetY;   /* :49-51 */
      } else if (algorithm == 2) {   /* :347-348 */
        inside = (i >= (nx / 2 - p->tipOffsetX)) && (i < (nx / 2 + p->tipOffsetX)) &&
                 (j >= (ny / 2 - p->tipOffsetY)) && (j < (ny / 2 + p->tipOffsetY));
      } else {
        inside = (i >= 1) && (i < (nx - 2)) && (j >= 1) && (j < (ny - 2));   /* :126 */
      }
      if (!inside) continue;
      double x1 = g_present[s0], x2 = g_present[sx], x4 = g_present[sy], x3 = g_present[sxy];
      double y1 = g_past[s0], y2 = g_past[sx], y4 = g_past[sy], y3 = g_past[sxy];
      if (algorithm == 1) {   /* :140-200 */
        double x3y1 = x3 * y1, x4y1 = x4 * y1, x3y2 = x3 * y2, x4y2 = x4 * y2;
        double x1y3 = x1 * y3, x2y3 = x2 * y3, x1y4 = x1 * y4, x2y4 = x2 * y4;
        double x2y1 = x2 * y1, x1y2 = x1 * y2, x4y3 = x4 * y3, x3y4 = x3 * y4;
        double den1 = 2.0 * (x3y1 - x4y1 - x3y2 + x4y2 - x1y3 + x2y3 + x1y4 - x2y4);
        double den2 = 2.0 * (x2y1 - x3y1 - x1y2 + x4y2 + x1y3 - x4y3 - x2y4 + x3y4);
        double ctn1 = x1 - x2 + x3 - x4 - y1 + y2 - y3 + y4;
        double ctn2 = x3y1 - 2.0 * x4y1 + x4y2 - x1y3 + 2.0 * x1y4 - x2y4;
        double disc = sqrt(4.0 * (x3y1 - x3y2 - x4y1 + x4y2 - x1y3 + x1y4 + x2y3 - x2y4)
                               * (x4y1 - x1y4 + Uth * (x1 - x4 - y1 + y4))
                           + (-ctn2 + Uth * ctn1) * (-ctn2 + Uth * ctn1));
        double px = ctn2 - Uth * ctn1;
        double py = Uth * ctn1 - x3y1 + x4y2 + x1y3 - x2y4 + 2.0 * (x2y1 - x1y2);
        int ok = p->solidSwitch ? 1 : (disc >= 0.0);   /* :185 vs :108 */
        float tx = (float)((px + disc) / den1), ty = (float)((py + disc) / den2);
        if (ok) tip_record(p, i, j, tx, ty, pTime, g_present, tip_plot, tip_count, tip_vector, capacity);
        tx = (float)((px - disc) / den1); ty = (float)((py - disc) / den2);
        if (ok) tip_record(p, i, j, tx, ty, pTime, g_present, tip_plot, tip_count, tip_vector, capacity);
      } else if (algorithm == 2) {   /* Newton, :302-341 */
        double s = 0.5, t = 0.5;
        for (int k = 0; k < 4; k++) {
          double r1 = x1 * (1.0 - s) * (1.0 - t) + x2 * s * (1.0 - t) + x3 * s * t + x4 * (1.0 - s) * t - Uth;
          double r2 = y1 * (1.0 - s) * (1.0 - t) + y2 * s * (1.0 - t) + y3 * s * t + y4 * (1.0 - s) * t - Uth;
          double J11 = -x1 * (1.0 - t) + x2 * (1.0 - t) + x3 * t - x4 * t;
          double J21 = -y1 * (1.0 - t) + y2 * (1.0 - t) + y3 * t - y4 * t;
          double J12 = -x1 * (1.0 - s) - x2 * s + x3 * s + x4 * (1.0 - s);
          double J22 = -y1 * (1.0 - s) - y2 * s + y3 * s + y4 * (1.0 - s);
          double detJ = J11 * J22 - J12 * J21;
          if (!equals_tol(detJ, 0.0, 1e-14)) {
            double s_new = s - (J22 * r1 - J12 * r2) / detJ;
            double t_new = t - (-J21 * r1 + J11 * r2) / detJ;
            s = fmin(fmax(s_new, 0.0), 1.0);
            t = fmin(fmax(t_new, 0.0), 1.0);
          } else { s = -1.0; t = -1.0; }
        }
        int in01 = (s >= 0.0) && (s <= 1.0) && (t >= 0.0) && (t <= 1.0);
        double u1 = in01 ? x1 * (1 - s) * (1.0 - t) + x2 * s * (1.0 - t) + x3 * s * t + x4 * (1.0 - s) * t : 0.0;
        double u2 = in01 ? y1 * (1 - s) * (1.0 - t) + y2 * s * (1.0 - t) + y3 * s * t + y4 * (1.0 - s) * t : 0.0;
        if (equals_tol(u1, Uth, 1e-15) && equals_tol(u2, Uth, 1e-15))
          tip_record(p, i, j, (float)s, (float)t, pTime, g_present, tip_plot, tip_count, tip_vector, capacity);
      } else {
        return YH_ERR_INVALID_ARG;
      }
    }
  }
  return (*tip_count > capacity) ? YH_ERR_CAPACITY : YH_OK;
}

/* ------------------------------------------------------------------------------------
 * Symmetry reduction: slice_kernel + helpers, symmetryReduction.cu:72-262
 * ---------------------------------------------------------------------------------- */
static void disc_centre(const yh_params *p, int tip_count, const yh_tip *tv, int count,
                        int *cx, int *cy) {
  /* symmetryReduction.cu:98-104 / integralTrapz.cu:42-48; empty list keeps (tipx0,tipy0) */
  float fx = p->tipx0, fy = p->tipy0;
  if (count != 0 && tip_count > 0 && tv) { fx = tv[tip_count - 1].x; fy = tv[tip_count - 1].y; }
  *cx = (int)rintf(fx - (float)(p->nx / 2));   /* __float2int_rn */
  *cy = (int)rintf(fy - (float)(p->ny / 2));
}

static inline double fb2x(const yh_params *p, const double *f, int i, int j, int C, int E, int W,
                          const double *advx) {   /* convFB2ndOX :242-251 */
  int WW = I2D(p->nx, mir(i - 2, p->nx), j), EE = I2D(p->nx, mir(i + 2, p->nx), j);
  return (advx[C] > 0.0) ? (-3.0 * f[C] + 4.0 * f[E] - f[EE]) * p->invdx
                         : (3.0 * f[C] - 4.0 * f[W] + f[WW]) * p->invdx;
}
static inline double fb2y(const yh_params *p, const double *f, int i, int j, int C, int N, int S,
                          const double *advy) {   /* convFB2ndOY :253-262 */
  int SS = I2D(p->nx, i, mir(j - 2, p->ny)), NN = I2D(p->nx, i, mir(j + 2, p->ny));
  return (advy[C] > 0.0) ? (-3.0 * f[C] + 4.0 * f[N] - f[NN]) * p->invdy
                         : (3.0 * f[C] - 4.0 * f[S] + f[SS]) * p->invdy;
}
static inline double cen2x(const yh_params *p, const double *f, int i, int j, int E, int W) {
  int WW = I2D(p->nx, mir(i - 2, p->nx), j), EE = I2D(p->nx, mir(i + 2, p->nx), j);   /* :224-231 */
  return (f[EE] - 8.0 * f[E] + 8.0 * f[W] - f[WW]) * p->invdx * (1.0 / 6.0);
}
static inline double cen2y(const yh_params *p, const double *f, int i, int j, int N, int S) {
  int SS = I2D(p->nx, i, mir(j - 2, p->ny)), NN = I2D(p->nx, i, mir(j + 2, p->ny));   /* :233-240 */
  return (f[NN] - 8.0 * f[N] + 8.0 * f[S] - f[SS]) * p->invdy * (1.0 / 6.0);
}

/* the 12 tangent values of one cell; returns sc */
static int slice_cell(const yh_params *p, const double *gu, const double *gv,
                      const double *advx, const double *advy, int scheme, int cx, int cy,
                      int i, int j, double s[6], double s0[6]) {
  const int nx = p->nx, ny = p->ny;
  const int i2d = I2D(nx, i, j);
  int ic = i - nx / 2, jc = j - ny / 2;
  int sc = ((ic - cx) * (ic - cx) + (jc - cy) * (jc - cy)) < p->tipOffsetX * p->tipOffsetY;
  double x = (double)(i2d % nx);
  double y = (double)floorf((float)((i2d / nx) % nx));   /* :109-110 */
  int S = I2D(nx, i, mir(j - 1, ny)), N = I2D(nx, i, mir(j + 1, ny));
  int W = I2D(nx, mir(i - 1, nx), j), E = I2D(nx, mir(i + 1, nx), j);
  const double hx = p->hx, hy = p->hy;
  int on = (scheme == 1) ? 1 : sc;
  s[0] = on ? fb2x(p, gu, i, j, i2d, E, W, advx) : 0.0;   /* ux */
  s[1] = on ? fb2y(p, gu, i, j, i2d, N, S, advy) : 0.0;   /* uy */
  s[3] = on ? fb2x(p, gv, i, j, i2d, E, W, advx) : 0.0;   /* vx */
  s[4] = on ? fb2y(p, gv, i, j, i2d, N, S, advy) : 0.0;   /* vy */
  s[2] = on ? hx * x * s[1] - hy * y * s[0] : 0.0;        /* ut */
  s[5] = on ? hx * x * s[4] - hy * y * s[3] : 0.0;        /* vt */
  if (scheme == 1) {   /* :151-157 */
    s0[0] = s[0]; s0[1] = s[1]; s0[3] = s[3]; s0[4] = s[4];
    s0[2] = hx * x * s[1] - hy * y * s[0];
    s0[5] = hx * x * s[4] - hy * y * s[3];
  } else {             /* :191-202 */
    s0[0] = sc ? cen2x(p, gu, i, j, E, W) : 0.0;
    s0[1] = sc ? cen2y(p, gu, i, j, N, S) : 0.0;
    s0[3] = sc ? cen2x(p, gv, i, j, E, W) : 0.0;
    s0[4] = sc ? cen2y(p, gv, i, j, N, S) : 0.0;
    s0[2] = sc ? hx * x * s0[1] - hy * y * s0[0] : 0.0;
    s0[5] = sc ? hx * x * s0[4] - hy * y * s0[3] : 0.0;
  }
  return sc;
}

int yho_slice(const yh_params *p, const double *u, const double *v,
              double *const slice[6], double *const slice0[6],
              int reduce_sym, int reduce_sym_start,
              const double *adv_x, const double *adv_y, int scheme,
              int tip_count, const yh_tip *tip_vector, int count) {
  if (!p || !u || !v || !slice || !adv_x || !adv_y) return YH_ERR_INVALID_ARG;
  if (scheme != 1 && scheme != 2) return YH_ERR_INVALID_ARG;
  if (!reduce_sym) return YH_OK;   /* :127,:169 nothing written */
  int cx, cy;
  disc_centre(p, tip_count, tip_vector, count, &cx, &cy);
#pragma omp parallel for schedule(static)
  for (int j = 0; j < p->ny; j++)
    for (int i = 0; i < p->nx; i++) {
      double s[6], s0[6];
      slice_cell(p, u, v, adv_x, adv_y, scheme, cx, cy, i, j, s, s0);
      size_t c = (size_t)i + (size_t)p->nx * j;
      for (int k = 0; k < 6; k++) slice[k][c] = s[k];
      if (reduce_sym_start && slice0)
        for (int k = 0; k < 6; k++) slice0[k][c] = s0[k];
    }
  return YH_OK;
}

/*
 * Canonical summation order of the 12 integrals (the reference's is nondeterministic:
 * 256 blocks atomicAdd(double), integralTrapz.cu:78-80):
 *   row sum   : 256 accumulators (8 warps x 32 lanes), accumulator a takes i = a, a+256, ...
 *               ascending, each term 4.0*(f*g + h*w) inside the disc (:55-57; the scb / 2.0
 *               branches multiply 0); the 32 lanes of a warp combined by the xor-butterfly
 *               16,8,4,2,1 (x = x + partner); the 8 warp sums added in ascending warp index;
 *   total     : row slots w = 0 .. 2R (slot w = grid row cy + ny/2 - R + w, R = ceil(sqrt(tipOffsetX*
 *               tipOffsetY)); rows outside the grid hold +0.0): 32 partial sums, partial l takes
 *               w = l, l+32, ... ascending; partials combined by the xor-butterfly 16..1;
 *               result = (0.25*hx*hy) * total   (:79).
 */
#define YHO_ROW_WARPS 8
typedef struct { double a[12]; } acc12;

static void integrals_rowsum(double acc[YHO_ROW_WARPS * 32][12], double out[12]) {
  for (int k = 0; k < 12; k++) out[k] = 0.0;
  for (int w = 0; w < YHO_ROW_WARPS; w++) {
    double (*lane)[12] = acc + 32 * w;
    for (int m = 16; m >= 1; m >>= 1) {
      double t[32][12];
      for (int l = 0; l < 32; l++)
        for (int k = 0; k < 12; k++) t[l][k] = lane[l][k] + lane[l ^ m][k];
      memcpy(lane, t, sizeof(t));
    }
    for (int k = 0; k < 12; k++) out[k] = (w == 0) ? lane[0][k] : out[k] + lane[0][k];
  }
}

/* pairs: Int[3a+b] = <slice0.a , slice.b>, Int[9+a] = <slice0.a , velTan>, a,b in {x,y,t} */
static inline void integrand12(const double s[6], const double s0[6], double vtu, double vtv,
                               double out[12]) {
  for (int a = 0; a < 3; a++) {
    for (int b = 0; b < 3; b++)
      out[3 * a + b] = 4.0 * (s0[a] * s[b] + s0[a + 3] * s[b + 3]);
    out[9 + a] = 4.0 * (s0[a] * vtu + s0[a + 3] * vtv);
  }
}

typedef void (*cell12_fn)(void *ctx, int i, int j, int *sc, double s[6], double s0[6]);

static int integrals_generic(const yh_params *p, cell12_fn fn, void *ctx, int cy,
                             const double *vtu, const double *vtv, double *integrals) {
  const int nx = p->nx, ny = p->ny;
  double *rows = (double *)calloc((size_t)ny * 12, sizeof(double));
  if (!rows) return YH_ERR_INVALID_ARG;
#pragma omp parallel for schedule(static)
  for (int j = 0; j < ny; j++) {
    double lane[YHO_ROW_WARPS * 32][12];
    memset(lane, 0, sizeof(lane));
    for (int i = 0; i < nx; i++) {
      double s[6], s0[6], t12[12];
      int sc;
      fn(ctx, i, j, &sc, s, s0);
      if (!sc) continue;   /* adds exactly +0.0 in the kernel; skipped terms do not change sums
                              because every accumulator starts at +0.0 and x + 0.0 == x */
      size_t c = (size_t)i + (size_t)nx * j;
      integrand12(s, s0, vtu[c], vtv[c], t12);
      for (int k = 0; k < 12; k++) lane[i % (YHO_ROW_WARPS * 32)][k] += t12[k];
    }
    integrals_rowsum(lane, rows + (size_t)j * 12);
  }
  int R = (int)ceil(sqrt((double)((long long)p->tipOffsetX * p->tipOffsetY)));
  if (R > ny) R = ny;
  double part[32][12];
  memset(part, 0, sizeof(part));
  for (int w = 0; w <= 2 * R; w++) {
    const int j = cy + ny / 2 - R + w;
    if (j < 0 || j >= ny) continue;   /* +0.0 */
    for (int k = 0; k < 12; k++) part[w & 31][k] += rows[(size_t)j * 12 + k];
  }
  for (int m = 16; m >= 1; m >>= 1) {
    double t[32][12];
    for (int l = 0; l < 32; l++)
      for (int k = 0; k < 12; k++) t[l][k] = part[l][k] + part[l ^ m][k];
    memcpy(part, t, sizeof(t));
  }
  for (int k = 0; k < 12; k++) integrals[k] = 0.25 * p->hx * p->hy * part[0][k];
  free(rows);
  return YH_OK;
}

typedef struct {
  const yh_params *p; const double *const *slice; const double *const *slice0; int cx, cy;
} trapz_ctx;
static void trapz_cell(void *vctx, int i, int j, int *sc, double s[6], double s0[6]) {
  trapz_ctx *c = (trapz_ctx *)vctx;
  const yh_params *p = c->p;
  int ic = i - p->nx / 2, jc = j - p->ny / 2;
  *sc = ((ic - c->cx) * (ic - c->cx) + (jc - c->cy) * (jc - c->cy)) < p->tipOffsetX * p->tipOffsetY;
  size_t idx = (size_t)i + (size_t)p->nx * j;
  for (int k = 0; k < 6; k++) { s[k] = c->slice[k][idx]; s0[k] = c->slice0[k][idx]; }
}

int yho_trapz(const yh_params *p, const double *const slice[6], const double *const slice0[6],
              const double *velTan_u, const double *velTan_v, double *integrals,
              int tip_count, const yh_tip *tip_vector, int count) {
  if (!p || !slice || !slice0 || !velTan_u || !velTan_v || !integrals) return YH_ERR_INVALID_ARG;
  trapz_ctx c = {p, slice, slice0, 0, 0};
  disc_centre(p, tip_count, tip_vector, count, &c.cx, &c.cy);
  return integrals_generic(p, trapz_cell, &c, c.cy, velTan_u, velTan_v, integrals);
}

typedef struct {
  const yh_params *p; const double *u, *v, *ax, *ay; int cx, cy;
} sri_ctx;
static void sri_cell(void *vctx, int i, int j, int *sc, double s[6], double s0[6]) {
  sri_ctx *c = (sri_ctx *)vctx;
  *sc = slice_cell(c->p, c->u, c->v, c->ax, c->ay, 2, c->cx, c->cy, i, j, s, s0);
}

int yho_sr_integrals(const yh_params *p, const double *u, const double *v,
                     const double *velTan_u, const double *velTan_v,
                     const double *adv_x, const double *adv_y, double *integrals,
                     int tip_count, const yh_tip *tip_vector, int count) {
  if (!p || !u || !v || !velTan_u || !velTan_v || !adv_x || !adv_y || !integrals)
    return YH_ERR_INVALID_ARG;
  sri_ctx c = {p, u, v, adv_x, adv_y, 0, 0};
  disc_centre(p, tip_count, tip_vector, count, &c.cx, &c.cy);
  return integrals_generic(p, sri_cell, &c, c.cy, velTan_u, velTan_v, integrals);
}

/* symmetryReduction.cu:386-416 */
int yho_solve_matrix(const double c_in[3], const double phi[3], const double Int[12],
                     double c_out[3]) {
  (void)c_in;
  double a1, a2, a3, b1, b2, b3, C1, C2, C3, d1, d2, d3;
  double b2p, b3p, c2p, c3p, c3pp, d2p, d3p, d3pp, x1, x2, x3;
  const double pt = phi[2];
  a1 = Int[0] * cos(pt) + Int[1] * sin(pt); a2 = Int[1] * cos(pt) - Int[0] * sin(pt); a3 = Int[2];
  b1 = Int[3] * cos(pt) + Int[4] * sin(pt); b2 = Int[4] * cos(pt) - Int[3] * sin(pt); b3 = Int[5];
  C1 = Int[6] * cos(pt) + Int[7] * sin(pt); C2 = Int[7] * cos(pt) - Int[6] * sin(pt); C3 = Int[8];
  d1 = Int[9]; d2 = Int[10]; d3 = Int[11];
  b2p = a1 / b1 * b2 - a2;
  b3p = a1 / b1 * b3 - a3;
  d2p = a1 / b1 * d2 - d1;
  c2p = a1 / C1 * C2 - a2;
  c3p = a1 / C1 * C3 - a3;
  d3p = a1 / C1 * d3 - d1;
  c3pp = b2p / c2p * c3p - b3p;
  d3pp = b2p / c2p * d3p - d2p;
  x3 = d3pp / c3pp;
  x2 = (d2p - b3p * x3) / b2p;
  x1 = (d1 - a2 * x2 - a3 * x3) / a1;
  c_out[0] = x1; c_out[1] = x2; c_out[2] = x3;
  return YH_OK;
}

/* symmetryReduction.cu:20-62 */
int yho_cxy_field(const yh_params *p, double *adv_x, double *adv_y,
                  const double c[3], const double phi[3], const uint8_t *solid) {
  if (!p || !adv_x || !adv_y || !c || !phi) return YH_ERR_INVALID_ARG;
  if (p->solidSwitch && !solid) return YH_ERR_INVALID_ARG;
  const int nx = p->nx, ny = p->ny;
  const double cs = cos(phi[2]), sn = sin(phi[2]);
#pragma omp parallel for schedule(static)
  for (int j = 0; j < ny; j++)
    for (int i = 0; i < nx; i++) {
      int i2d = I2D(nx, i, j);
      double x = (double)(i2d % nx);
      double y = (double)floorf((float)((i2d / nx) % nx));
      double ax = p->hy * y * c[2] - c[0] * cs + c[1] * sn;
      double ay = -p->hx * x * c[2] - c[0] * sn - c[1] * cs;
      int sc = p->solidSwitch ? (solid[i2d] != 0) : 1;
      adv_x[i2d] = sc ? ax : 0.0;
      adv_y[i2d] = sc ? ay : 0.0;
    }
  return YH_OK;
}

/* ------------------------------------------------------------------------------------
 * BFECC advection, advFDBFECC.cu:18-353, three synchronous sweeps
 * ---------------------------------------------------------------------------------- */
static inline int sgn(double x) { int t = x < 0.0 ? -1 : 0; return x > 0.0 ? 1 : t; }   /* helper_functions.cu:151 */

typedef struct {
  int W, E, S, N;                       /* forward */
  int i2dW, W2, i2dE, E2, i2dS, S2, i2dN, N2;   /* backward */
  int sc;
} bfecc_idx;

static inline int clampmir(const yh_params *p, int i, int j) {
  return I2D(p->nx, mir(i, p->nx), mir(j, p->ny));
}

static void bfecc_neumann_idx(const yh_params *p, const uint8_t *solid, int i, int j, bfecc_idx *x) {
  const int nx = p->nx, ny = p->ny;
  if (p->solidSwitch) {   /* :44-66; out-of-domain mask = non-tissue, indices mirrored (B5) */
    int sc = solid[I2D(nx, i, j)] != 0;
    int sw = solid_at(solid, nx, ny, i - 1, j), se = solid_at(solid, nx, ny, i + 1, j);
    int sn = solid_at(solid, nx, ny, i, j - 1), ss = solid_at(solid, nx, ny, i, j + 1);
    int C = I2D(nx, i, j);
    x->sc = sc;
    x->W = sc ? (sw ? clampmir(p, i - 1, j) : clampmir(p, i + 1, j)) : C;
    x->E = sc ? (se ? clampmir(p, i + 1, j) : clampmir(p, i - 1, j)) : C;
    x->S = sc ? (ss ? clampmir(p, i, j - 1) : clampmir(p, i, j + 1)) : C;
    x->N = sc ? (sn ? clampmir(p, i, j + 1) : clampmir(p, i, j - 1)) : C;
    x->i2dW = sc ? (sw ? C : clampmir(p, i + 1, j)) : C;
    x->W2 = sc ? (sw ? clampmir(p, i - 1, j) : C) : C;
    x->i2dE = sc ? (se ? C : clampmir(p, i - 1, j)) : C;
    x->E2 = sc ? (se ? clampmir(p, i + 1, j) : C) : C;
    x->i2dS = sc ? (ss ? C : clampmir(p, i, j + 1)) : C;
    x->S2 = sc ? (ss ? clampmir(p, i, j - 1) : C) : C;
    x->i2dN = sc ? (sn ? C : clampmir(p, i, j - 1)) : C;
    x->N2 = sc ? (sn ? clampmir(p, i, j + 1) : C) : C;
  } else {   /* :112-126 */
    int C = I2D(nx, i, j);
    x->sc = 1;
    x->W = (i > 0) ? I2D(nx, i - 1, j) : I2D(nx, i + 1, j);
    x->E = (i < (nx - 1)) ? I2D(nx, i + 1, j) : I2D(nx, i - 1, j);
    x->S = (j > 0) ? I2D(nx, i, j - 1) : I2D(nx, i, j + 1);
    x->N = (j < (ny - 1)) ? I2D(nx, i, j + 1) : I2D(nx, i, j - 1);
    x->i2dW = (i > 0) ? C : I2D(nx, i + 1, j);
    x->W2 = (i > 0) ? I2D(nx, i - 1, j) : C;
    x->i2dE = (i < (nx - 1)) ? C : I2D(nx, i - 1, j);
    x->E2 = (i < (nx - 1)) ? I2D(nx, i + 1, j) : C;
    x->i2dS = (j > 0) ? C : I2D(nx, i, j + 1);
    x->S2 = (j > 0) ? I2D(nx, i, j - 1) : C;
    x->i2dN = (j < (ny - 1)) ? C : I2D(nx, i, j - 1);
    x->N2 = (j < (ny - 1)) ? I2D(nx, i, j + 1) : C;
  }
}

/* One field through the Neumann BFECC pipeline (both solid and square share the form). */
static void bfecc_neumann_field(const yh_params *p, const uint8_t *solid, const double *g,
                                const double *advx, const double *advy, double *out,
                                double *uf, double *ub_unused, double *ue) {
  (void)ub_unused;
  const int nx = p->nx, ny = p->ny;
  const double tc = p->tc;
#pragma omp parallel for schedule(static)
  for (int j = 0; j < ny; j++)
    for (int i = 0; i < nx; i++) {
      int C = I2D(nx, i, j);
      bfecc_idx x; bfecc_neumann_idx(p, solid, i, j, &x);
      double cx = -advx[C], cy = -advy[C];
      double Rx = sgn(cx) * cx * p->dt / p->hx, Ry = sgn(cy) * cy * p->dt / p->hy;   /* :33-34 */
      double FDx = cx > 0.0 ? g[C] - g[x.W] : g[C] - g[x.E];
      double FDy = cy > 0.0 ? g[C] - g[x.S] : g[C] - g[x.N];
      uf[C] = g[C] - tc * (Rx * FDx + Ry * FDy);   /* :131 */
    }
#pragma omp parallel for schedule(static)
  for (int j = 0; j < ny; j++)
    for (int i = 0; i < nx; i++) {
      int C = I2D(nx, i, j);
      bfecc_idx x; bfecc_neumann_idx(p, solid, i, j, &x);
      double cx = -advx[C], cy = -advy[C];
      double Rx = sgn(cx) * cx * p->dt / p->hx, Ry = sgn(cy) * cy * p->dt / p->hy;
      double FDx = cx > 0.0 ? uf[x.i2dE] - uf[x.E2] : uf[x.i2dW] - uf[x.W2];   /* :134-135 */
      double FDy = cy > 0.0 ? uf[x.i2dN] - uf[x.N2] : uf[x.i2dS] - uf[x.S2];
      double ubv = uf[C] - tc * (Rx * FDx + Ry * FDy);   /* :137 */
      ue[C] = g[C] - 0.5 * (ubv - g[C]);                 /* :139 */
    }
#pragma omp parallel for schedule(static)
  for (int j = 0; j < ny; j++)
    for (int i = 0; i < nx; i++) {
      int C = I2D(nx, i, j);
      bfecc_idx x; bfecc_neumann_idx(p, solid, i, j, &x);
      double cx = -advx[C], cy = -advy[C];
      double Rx = sgn(cx) * cx * p->dt / p->hx, Ry = sgn(cy) * cy * p->dt / p->hy;
      double FDx = cx > 0.0 ? ue[C] - ue[x.W] : ue[C] - ue[x.E];   /* :141-142 */
      double FDy = cy > 0.0 ? ue[C] - ue[x.S] : ue[C] - ue[x.N];
      double r = ue[C] - tc * (Rx * FDx + Ry * FDy);               /* :144 */
      out[C] = x.sc ? r : 0.0;                                     /* :83 */
    }
}

/* Dirichlet branches, :167-347.  is_u selects the shipped sign typo of :287 (square, u only). */
static void bfecc_dirichlet_field(const yh_params *p, const uint8_t *solid, const double *g,
                                  const double *advx, const double *advy, double *out,
                                  double *uf, double *ue, int is_u) {
  const int nx = p->nx, ny = p->ny;
  const double tc = p->tc, bv = p->boundaryVal;
  const int so = p->solidSwitch;
#define MASKS                                                                          \
  int sc = so ? (solid[C] != 0) : 1;                                                   \
  int sw = so ? solid_at(solid, nx, ny, i - 1, j) : (i > 0);                           \
  int se = so ? solid_at(solid, nx, ny, i + 1, j) : (i < (nx - 1));                    \
  int sn = so ? solid_at(solid, nx, ny, i, j - 1) : (j > 0);       /* "sn" = j-1 */    \
  int ss = so ? solid_at(solid, nx, ny, i, j + 1) : (j < (ny - 1)); /* "ss" = j+1 */   \
  double cx = -advx[C], cy = -advy[C];                                                 \
  double Rx = sgn(cx) * cx * p->dt / p->hx, Ry = sgn(cy) * cy * p->dt / p->hy;
  /* In the solid branch the reference reads S from (i,j-1) gated by ss (mask at j+1) and N
   * from (i,j+1) gated by sn (mask at j-1), :182-183; the square branch gates by the index
   * itself, :268-269.  Both are reproduced. */
#pragma omp parallel for schedule(static)
  for (int j = 0; j < ny; j++)
    for (int i = 0; i < nx; i++) {
      int C = I2D(nx, i, j);
      MASKS
      double u = sc ? g[C] : 0.0;
      double W, E, S, N;
      if (so) {
        W = sc && sw ? g[I2D(nx, i - 1, j)] : (sc && !sw ? bv : 0.0);
        E = sc && se ? g[I2D(nx, i + 1, j)] : (sc && !se ? bv : 0.0);
        S = sc && ss ? g[clampmir(p, i, j - 1)] : (sc && !ss ? bv : 0.0);
        N = sc && sn ? g[clampmir(p, i, j + 1)] : (sc && !sn ? bv : 0.0);
      } else {
        W = sw ? g[I2D(nx, i - 1, j)] : bv; E = se ? g[I2D(nx, i + 1, j)] : bv;
        S = sn ? g[I2D(nx, i, j - 1)] : bv; N = ss ? g[I2D(nx, i, j + 1)] : bv;
      }
      double FDx = cx > 0.0 ? u - W : u - E, FDy = cy > 0.0 ? u - S : u - N;
      uf[C] = u - tc * (Rx * FDx + Ry * FDy);
    }
#pragma omp parallel for schedule(static)
  for (int j = 0; j < ny; j++)
    for (int i = 0; i < nx; i++) {
      int C = I2D(nx, i, j);
      MASKS
      double u = sc ? g[C] : 0.0;
      double uuf = sc ? uf[C] : 0.0;
      double W, E, S, N;
      if (so) {
        W = sc && sw ? uf[I2D(nx, i - 1, j)] : (sc && !sw ? uf[C] : 0.0);
        E = sc && se ? uf[I2D(nx, i + 1, j)] : (sc && !se ? uf[C] : 0.0);
        S = sc && ss ? uf[clampmir(p, i, j - 1)] : (sc && !ss ? uf[C] : 0.0);
        N = sc && sn ? uf[clampmir(p, i, j + 1)] : (sc && !sn ? uf[C] : 0.0);
      } else {
        W = sw ? uf[I2D(nx, i - 1, j)] : uf[C]; E = se ? uf[I2D(nx, i + 1, j)] : uf[C];
        S = sn ? uf[I2D(nx, i, j - 1)] : uf[C]; N = ss ? uf[I2D(nx, i, j + 1)] : uf[C];
      }
      double FDx = cx > 0.0 ? uuf - E : uuf - W, FDy = cy > 0.0 ? uuf - N : uuf - S;
      double ubv = (!so && is_u) ? uuf - tc * (Rx * FDx - Ry * FDy)    /* :287 as shipped */
                                 : uuf - tc * (Rx * FDx + Ry * FDy);
      ue[C] = u - 0.5 * (ubv - u);
    }
#pragma omp parallel for schedule(static)
  for (int j = 0; j < ny; j++)
    for (int i = 0; i < nx; i++) {
      int C = I2D(nx, i, j);
      MASKS
      double uue = sc ? ue[C] : 0.0;
      double W, E, S, N;
      if (so) {
        W = sc && sw ? ue[I2D(nx, i - 1, j)] : (sc && !sw ? bv : 0.0);
        E = sc && se ? ue[I2D(nx, i + 1, j)] : (sc && !se ? bv : 0.0);
        S = sc && ss ? ue[clampmir(p, i, j - 1)] : (sc && !ss ? bv : 0.0);
        N = sc && sn ? ue[clampmir(p, i, j + 1)] : (sc && !sn ? bv : 0.0);
      } else {
        W = sw ? ue[I2D(nx, i - 1, j)] : bv; E = se ? ue[I2D(nx, i + 1, j)] : bv;
        S = sn ? ue[I2D(nx, i, j - 1)] : bv; N = ss ? ue[I2D(nx, i, j + 1)] : bv;
      }
      double FDx = cx > 0.0 ? uue - W : uue - E, FDy = cy > 0.0 ? uue - S : uue - N;
      double r = uue - tc * (Rx * FDx + Ry * FDy);
      out[C] = sc ? r : 0.0;
    }
#undef MASKS
}

int yho_advect_bfecc(const yh_params *p, const double *u_in, const double *v_in,
                     double *u_out, double *v_out,
                     const double *adv_x, const double *adv_y, const uint8_t *solid) {
  if (!p || !u_in || !v_in || !u_out || !v_out || !adv_x || !adv_y) return YH_ERR_INVALID_ARG;
  if (p->solidSwitch && !solid) return YH_ERR_INVALID_ARG;
  const size_t n = (size_t)p->nx * p->ny;
  double *uf = (double *)malloc(n * sizeof(double)), *ue = (double *)malloc(n * sizeof(double));
  if (!uf || !ue) return YH_ERR_INVALID_ARG;
  if (p->neumannBC) {
    bfecc_neumann_field(p, solid, u_in, adv_x, adv_y, u_out, uf, NULL, ue);
    bfecc_neumann_field(p, solid, v_in, adv_x, adv_y, v_out, uf, NULL, ue);
  } else {
    bfecc_dirichlet_field(p, solid, u_in, adv_x, adv_y, u_out, uf, ue, 1);
    bfecc_dirichlet_field(p, solid, v_in, adv_x, adv_y, v_out, uf, ue, 0);
  }
  free(uf); free(ue);
  return YH_OK;
}

/* ------------------------------------------------------------------------------------
 * sAPD_kernel, spaceAPD.cu:278-374
 * ---------------------------------------------------------------------------------- */
int yho_sapd(const yh_params *p, int count, const double *uold, const double *unew,
             double *APD1, double *APD2, double *sAPD, double *dAPD,
             double *back, double *front, uint8_t *first, const uint8_t *stimArea,
             int stimulate) {
  if (!p || !uold || !unew || !APD1 || !APD2 || !sAPD || !back || !front || !first)
    return YH_ERR_INVALID_ARG;
  if (stimulate && (!stimArea || !dAPD)) return YH_ERR_INVALID_ARG;
  const size_t n = (size_t)p->nx * p->ny;
  const double apdTh = 0.15, dt = p->dt;
#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < n; k++) {
    double uo = uold[k], un = unew[k];
    int sc = stimulate ? (stimArea[k] != 0) : 1;
    if ((un > apdTh) && (uo < apdTh) && sc) front[k] = dt * (count - (un - apdTh) / (un - uo));
    if ((un < apdTh) && (uo > apdTh) && sc) back[k] = dt * (count - (un - apdTh) / (un - uo));
    if ((back[k] > 0.0) && (front[k] > 0.0) && (first[k] == 0) && sc) {
      APD1[k] = back[k] - front[k]; front[k] = 0.0; back[k] = 0.0; first[k] = 1;
    }
    if ((back[k] > 0.0) && (front[k] > 0.0) && first[k] && sc) {
      APD2[k] = back[k] - front[k]; front[k] = 0.0; back[k] = 0.0; first[k] = 0;
    }
    if (stimulate) {
      double s = (APD1[k] - APD2[k] > 0.0) && sc ? 1.0 : -1.0;
      s *= (double)sc;
      sAPD[k] = s;
      double d = APD2[k];
      d *= (double)sc;
      dAPD[k] = d;
    } else {
      sAPD[k] = (APD1[k] - APD2[k] > 0.0) ? 1.0 : -1.0;
    }
  }
  return YH_OK;
}

/* ------------------------------------------------------------------------------------
 * Contours: countour_kernel modes 1-3, spaceAPD.cu:18-153 (+ countour_wrapper :256-276),
 * list in canonical order (ascending i + j*nx; mode 3: -conTh2 crossing first).
 * Reads past the end of the array (last row / last cell, :52-53) return the cell's own value.
 * ------------------------------------------------------------------------------------ */
static void contour_push(const yh_params *p, double ppx, double ppy, double pTime,
                         uint8_t *plot, int *count, yh_contour_pt *vec, int capacity) {
  yh_contour_pt d;
  d.x = (float)ppx; d.y = (float)ppy; d.t = (float)pTime;   /* make_float3, :66 */
  if (*count < capacity) vec[*count] = d;
  (*count)++;
  if (plot) {   /* plot_field, helper_functions.cu:45-51 */
    float fx = floorf(d.x), fy = floorf(d.y);
    if (fabsf(fx) < 1e9f && fabsf(fy) < 1e9f) {
      long long idx = (long long)fx + (long long)p->nx * (long long)fy;
      if (idx >= 0 && idx < (long long)p->nx * p->ny) plot[idx] = 1;
    }
  }
}

int yho_contour(const yh_params *p, const double *field1, const double *field2,
                uint8_t *contour_plot, const uint8_t *stimArea, int *contour_count,
                yh_contour_pt *contour_vector, int capacity, double pTime, int mode,
                double th1, double th2, double th3) {
  if (!p || !field2 || !contour_count || !contour_vector || mode < 1 || mode > 3) return YH_ERR_INVALID_ARG;
  if (mode != 1 && !field1) return YH_ERR_INVALID_ARG;
  const int nx = p->nx, ny = p->ny;
  const long long ncell = (long long)nx * ny;
  *contour_count = 0;                                            /* :258 */
  if (contour_plot) memset(contour_plot, 0, (size_t)ncell);      /* :268 */
  for (int j = 0; j < ny; j++) {
    for (int i = 0; i < nx; i++) {
      const long long c = I2D(nx, i, j);
      const int sc = stimArea ? stimArea[c] != 0 : 1;            /* :43 */
      const double f0 = field2[c];
      const double fx = (c + 1 < ncell) ? field2[c + 1] : f0;    /* I2D(nx_d,i+1,j), unclamped */
      const double fy = (c + nx < ncell) ? field2[c + nx] : f0;  /* I2D(nx_d,i,j+1) */
      if (mode == 1) {                                           /* :50-71 */
        double v0 = f0, v1x = fx, v1y = fy;
        double zpmx = i < (nx - 1) ? v0 * v1x : v0 * v0;
        double zpmy = j < (ny - 1) ? v0 * v1y : v0 * v0;
        if ((zpmx < 0.0) || ((zpmy < 0.0) && sc)) {              /* `a || b && sc` as written, :62 */
          double ppx = fabs(v0 - v1x) > 2.220446049250313e-16 ? i + v0 / (v0 - v1x) : i;
          double ppy = fabs(v0 - v1y) > 2.220446049250313e-16 ? j + v0 / (v0 - v1y) : j;
          contour_push(p, ppx, ppy, pTime, contour_plot, contour_count, contour_vector, capacity);
        }
      } else if (mode == 2) {                                    /* :74-95 */
        if ((field1[c] < th1) && sc) {
          double v0 = f0 - th2, v1x = fx - th2, v1y = fy - th2;
          double zpmx = i < (nx - 1) ? v0 * v1x : v0 * v0;
          double zpmy = j < (ny - 1) ? v0 * v1y : v0 * v0;
          if ((zpmx < 0.0) || (zpmy < 0.0))
            contour_push(p, i, j, pTime, contour_plot, contour_count, contour_vector, capacity);
        }
      } else {                                                   /* :97-141 */
        double V0 = f0 - th2, V1x = fx - th2, V1y = fy - th2;
        if ((field1[c] < th1) && (fabs(V0) < (th3 + 0.1)) && (fabs(V1x) < (th3 + 0.1)) &&
            (fabs(V1y) < (th3 + 0.1)) && sc) {
          double v0 = V0 - th2, v1x = V1x - th2, v1y = V1y - th2;
          double zpmx = i < (nx - 1) ? v0 * v1x : v0 * v0;
          double zpmy = j < (ny - 1) ? v0 * v1y : v0 * v0;
          if ((zpmx < 0.0) || (zpmy < 0.0))
            contour_push(p, i, j, pTime, contour_plot, contour_count, contour_vector, capacity);
          v0 = V0 + th3; v1x = V1x + th3; v1y = V1y + th3;
          zpmx = i < (nx - 1) ? v0 * v1x : v0 * v0;
          zpmy = j < (ny - 1) ? v0 * v1y : v0 * v0;
          if ((zpmx < 0.0) || (zpmy < 0.0))
            contour_push(p, i, j, pTime, contour_plot, contour_count, contour_vector, capacity);
        }
      }
    }
  }
  return YH_OK;
}

/* get_rgba_kernel, main.cu:1604-1631 (colour index clamped to the map) */
int yho_rgba(const yh_params *p, const double *field, uint32_t *plot_rgba, const uint32_t *cmap,
             int ncol, double vmin, double vmax, const uint8_t *lines) {
  if (!p || !field || !plot_rgba || !cmap || ncol <= 0) return YH_ERR_INVALID_ARG;
  const long long n = (long long)p->nx * p->ny;
  for (long long c = 0; c < n; c++) {
    double frac = (field[c] - vmin) / (vmax - vmin);
    int icol = (int)((float)frac * (float)ncol);
    icol = icol < 0 ? 0 : (icol >= ncol ? ncol - 1 : icol);
    plot_rgba[c] = (uint32_t)(lines ? !lines[c] : 1) * cmap[icol];
  }
  return YH_OK;
}
