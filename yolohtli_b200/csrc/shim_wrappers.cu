// shim_wrappers.cu -> libyolohtli_shim.so: the reference's launch API (hostPrototypes.h:22-57)
// with the reference's exact C++ signatures, forwarding to the C ABI of libyolohtli_b200.so.
// pitch / grid / block arguments are hints and are ignored (the new kernels choose their own
// launch shape); scalars come from yh_shim_configure() instead of __constant__ symbols.
// Wrappers are `void` like the reference's; a failure is printed, latched in
// yh_shim_last_status(), and -- like -DCUDA_ERROR_CHECK builds (common/CudaSafeCall.h:8-16) --
// terminates the process when YH_SHIM_ABORT_ON_ERROR is set in the environment.
#include <stdio.h>
#include <stdlib.h>

#include "../../include/yolohtli_abi.h"
#include "../../include/yolohtli_compat.h"

static yh_params g_p;
static bool g_configured = false;     // yh_shim_configure() was called: the explicit block wins
static int g_status = YH_OK;
static double g_con_th[3] = {0.8, 0.85, 0.7};
static double g_color[2] = {-0.1f, 1.1f};

// The reference's global run-parameter block (main.cu:37).  Weak: the shim also loads into programs that
// do not have it (then yh_shim_configure is the way in).
extern paramVar param __attribute__((weak));

// paramVar -> yh_params, field for field what main.cu:309-402 copies into the __constant__ symbols.
static void from_reference_param(void) {
  const paramVar &q = param;
  yh_params p;
  p.nx = q.nx; p.ny = q.ny; p.ny_global = q.ny; p.jg0 = 0;
  p.solidSwitch = q.solidSwitch; p.neumannBC = q.neumannBC; p.gateDiff = q.gateDiff; p.anisotropy = q.anisotropy;
  p.lap4 = q.lap4; p.timeIntOrder = q.timeIntOrder; p.tipGrad = q.tipGrad; p.tipAlgorithm = q.tipAlgorithm;
  p.tipOffsetX = q.tipOffsetX; p.tipOffsetY = q.tipOffsetY; p.tipx0 = q.tipx; p.tipy0 = q.tipy;   // main.cu:373-376
  p.dt = q.dt; p.hx = q.hx; p.hy = q.hy; p.Lx = q.Lx; p.Ly = q.Ly;
  p.rx = q.rx; p.ry = q.ry; p.rxy = q.rxy; p.rbx = q.rbx; p.rby = q.rby; p.rscale = q.rscale;
  p.qx4 = q.qx4; p.qy4 = q.qy4; p.fx4 = q.fx4; p.fy4 = q.fy4; p.invdx = q.invdx; p.invdy = q.invdy;
  p.tc = q.tc; p.alpha = q.alpha; p.beta = q.beta; p.gamma = q.gamma; p.delta = q.delta; p.eps = q.eps;
  p.mu = q.mu; p.theta = q.theta; p.boundaryVal = q.boundaryVal; p.Uth = q.Uth;
  g_p = p;
  g_con_th[0] = q.contourThresh1; g_con_th[1] = q.contourThresh2; g_con_th[2] = q.contourThresh3;
  if (q.minVarColor != q.maxVarColor) { g_color[0] = q.minVarColor; g_color[1] = q.maxVarColor; }
}

static void note(int rc, const char *what) {
  if (rc == YH_OK) return;
  g_status = rc;
  fprintf(stderr, "yolohtli shim: %s failed (%d): %s\n", what, rc, yh_last_error());
  if (getenv("YH_SHIM_ABORT_ON_ERROR")) exit(-1);
}
static bool ready(const char *what) {
  if (g_configured) return true;
  if (&param != nullptr) {      // the host is the reference's main.cu (or built like it): read its global,
    from_reference_param();     // at every call -- it is a few dozen scalars, and always current
    return true;
  }
  g_status = YH_ERR_INVALID_ARG;
  fprintf(stderr, "yolohtli shim: %s called before yh_shim_configure() in a program without the reference's "
                  "global `paramVar param`\n", what);
  if (getenv("YH_SHIM_ABORT_ON_ERROR")) exit(-1);
  return false;
}

extern "C" int yh_shim_configure(const yh_params *p) {
  if (!p) return YH_ERR_INVALID_ARG;
  g_p = *p;
  g_configured = true;
  g_status = YH_OK;
  return YH_OK;
}
extern "C" int yh_shim_last_status(void) { return g_status; }

void reactionDiffusion_wrapper(size_t, dim3, dim3, stateVar gOut_d, stateVar gIn_d, stateVar,
                               stateVar velTan, bool, bool *solid, bool, REAL *, bool stimLockMouse,
                               int2 point) {
  if (!ready("reactionDiffusion_wrapper")) return;
  // velTan is produced whenever gateDiff is on, as reactionDiffusion.cu:529-552 does; the J
  // scratch array and the stim/stimLock arguments are unused (the latter as shipped, :132-133).
  note(yh_rd_step(&g_p, gIn_d.u, gIn_d.v, gOut_d.u, gOut_d.v, g_p.gateDiff ? velTan.u : nullptr,
                  g_p.gateDiff ? velTan.v : nullptr, reinterpret_cast<const uint8_t *>(solid),
                  stimLockMouse, point.x, point.y, 0, g_p.ny, nullptr),
       "reactionDiffusion_wrapper");
}

void tip_wrapper(size_t, dim3, dim3, stateVar gOut_d, stateVar gIn_d, stateVar, REAL physicalTime,
                 int tipAlgorithm, bool, bool *tip_plot, int *tip_count, vec5dyn *tip_vector) {
  if (!ready("tip_wrapper")) return;
  // tipTracker.cu:586: kernel(g_past = gIn_d.u, g_present = gOut_d.u)
  note(yh_tip_track(&g_p, gIn_d.u, gOut_d.u, reinterpret_cast<uint8_t *>(tip_plot), tip_count,
                    reinterpret_cast<yh_tip *>(tip_vector), YH_TIPVECSIZE, physicalTime,
                    tipAlgorithm, nullptr),
       "tip_wrapper");
}

void slice_wrapper(size_t, dim3, dim3, stateVar g, sliceVar slice, sliceVar slice0, bool reduceSym,
                   bool reduceSymStart, advVar adv, int scheme, bool *, int *tip_count,
                   vec5dyn *tip_vector, int count) {
  if (!ready("slice_wrapper")) return;
  double *s[6] = {slice.ux, slice.uy, slice.ut, slice.vx, slice.vy, slice.vt};
  double *s0[6] = {slice0.ux, slice0.uy, slice0.ut, slice0.vx, slice0.vy, slice0.vt};
  note(yh_slice(&g_p, g.u, g.v, s, s0, reduceSym, reduceSymStart, adv.x, adv.y, scheme, tip_count,
                reinterpret_cast<const yh_tip *>(tip_vector), count, nullptr),
       "slice_wrapper");
}

void Cxy_field_wrapper(size_t, dim3, dim3, advVar adv, REAL3 c, REAL3 phi, bool *solid) {
  if (!ready("Cxy_field_wrapper")) return;
  const double cc[3] = {c.x, c.y, c.t}, ph[3] = {phi.x, phi.y, phi.t};
  note(yh_cxy_field(&g_p, adv.x, adv.y, cc, ph, reinterpret_cast<const uint8_t *>(solid), nullptr),
       "Cxy_field_wrapper");
}

void advFDBFECC_wrapper(size_t, dim3, dim3, stateVar gOut, stateVar gIn, advVar adv, stateVar,
                        stateVar, stateVar, bool *solid) {
  if (!ready("advFDBFECC_wrapper")) return;   // uf / ub / ue scratch arrays are not needed
  note(yh_advect_bfecc(&g_p, gIn.u, gIn.v, gOut.u, gOut.v, adv.x, adv.y,
                       reinterpret_cast<const uint8_t *>(solid), nullptr),
       "advFDBFECC_wrapper");
}

REAL3 solve_matrix(REAL3 c, REAL3 phi, REAL *Int) {
  const double ci[3] = {c.x, c.y, c.t}, ph[3] = {phi.x, phi.y, phi.t};
  double co[3] = {c.x, c.y, c.t};
  note(yh_solve_matrix(ci, ph, Int, co), "solve_matrix");
  REAL3 r = {co[0], co[1], co[2]};
  return r;
}

void trapz_wrapper(dim3, dim3, sliceVar slice, sliceVar slice0, stateVar velTan, REAL *integrals,
                   REAL *, int *tip_count, vec5dyn *tip_vector, int count) {
  if (!ready("trapz_wrapper")) return;
  const double *s[6] = {slice.ux, slice.uy, slice.ut, slice.vx, slice.vy, slice.vt};
  const double *s0[6] = {slice0.ux, slice0.uy, slice0.ut, slice0.vx, slice0.vy, slice0.vt};
  note(yh_trapz(&g_p, s, s0, velTan.u, velTan.v, integrals, tip_count,
                reinterpret_cast<const yh_tip *>(tip_vector), count, nullptr),
       "trapz_wrapper");
}

void singleCell_wrapper(size_t, dim3, dim3, stateVar gOut_d, int, REAL *pt_h, REAL *pt_d, int2 point) {
  if (!ready("singleCell_wrapper")) return;
  note(yh_probe(&g_p, gOut_d.u, gOut_d.v, pt_d, point.x, point.y, pt_h, nullptr), "singleCell_wrapper");
}

void sAPD_wrapper(size_t, dim3, dim3, int count, REAL *uold, REAL *unew, REAL *APD1, REAL *APD2,
                  REAL *sAPD, REAL *dAPD, REAL *back, REAL *front, bool *first, bool *stimArea,
                  bool stimulate) {
  if (!ready("sAPD_wrapper")) return;
  note(yh_sapd(&g_p, count, uold, unew, APD1, APD2, sAPD, dAPD, back, front,
               reinterpret_cast<uint8_t *>(first), reinterpret_cast<const uint8_t *>(stimArea),
               stimulate, nullptr),
       "sAPD_wrapper");
}

// conTh1_d..conTh3_d and minVarColor_d / maxVarColor_d are __constant__ symbols in the reference
// (main.cu:378-383, 366-369); defaults of parameterSetup (saveFiles.cu:196-217).
extern "C" int yh_shim_set_contour_thresholds(double th1, double th2, double th3) {
  g_con_th[0] = th1; g_con_th[1] = th2; g_con_th[2] = th3;
  return YH_OK;
}
extern "C" int yh_shim_set_color_range(double min_var, double max_var) {
  if (min_var == max_var) return YH_ERR_INVALID_ARG;
  g_color[0] = min_var; g_color[1] = max_var;
  return YH_OK;
}

void countour_wrapper(size_t, dim3, dim3, REAL *field1, REAL *field2, bool *contour_plot,
                      bool *stimArea, int *contour_count, float3 *contour_vector, float physicalTime,
                      int mode) {
  if (!ready("countour_wrapper")) return;
  // capacity: the reference allocates nx*ny float3 (main.cu:300)
  note(yh_contour(&g_p, field1, field2, reinterpret_cast<uint8_t *>(contour_plot),
                  reinterpret_cast<const uint8_t *>(stimArea), contour_count,
                  reinterpret_cast<yh_contour_pt *>(contour_vector), g_p.nx * g_p.ny, physicalTime,
                  mode, g_con_th[0], g_con_th[1], g_con_th[2], nullptr),
       "countour_wrapper");
}

void get_rgba_wrapper(size_t, dim3, dim3, int ncol, REAL *field, unsigned int *plot_rba_data,
                      unsigned int *cmap_rgba_data, bool *lines) {
  if (!ready("get_rgba_wrapper")) return;
  note(yh_rgba(&g_p, field, plot_rba_data, cmap_rgba_data, ncol, g_color[0], g_color[1],
               reinterpret_cast<const uint8_t *>(lines), nullptr),
       "get_rgba_wrapper");
}

void swapSoA(stateVar *A, stateVar *B) {
  stateVar t = *A;
  *A = *B;
  *B = t;
}
