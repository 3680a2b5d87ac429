/*
 * ref_harness.cu -- TEST INFRASTRUCTURE.  Headless driver for the UNMODIFIED reference
 * kernels of /root/reference (compiled in place by oracle/build_ref.sh; no reference source
 * is copied into this repository).  It supplies what main.cu supplies -- the __constant__
 * symbols (main.cu:40-51), `paramVar param` (main.cu:37) and a SOIL stub -- and exposes small
 * extern "C" entry points that move HOST arrays through the reference's own *_wrapper
 * functions (hostPrototypes.h:22-54) so tests can diff them against the oracle and the
 * new kernels, and bench.py --impl reference can time them.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's reference/cpu_baseline legs load the
 * resulting oracle/_ref/libyhref*.so.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cuda_runtime.h>
#include <malloc.h>

#include "typeDefinition.cuh"
#include "hostPrototypes.h"
#include "devicePrototypes.cuh"
#include "../include/yolohtli_abi.h"

// ---- what main.cu defines (main.cu:33-51) -------------------------------------------
size_t pitch;
paramVar param;
__constant__ int nx_d, ny_d;
__constant__ REAL dt_d, rx_d, ry_d, hx_d, hy_d, Lx_d, Ly_d, qx4_d, qy4_d, fx4_d, fy4_d;
__constant__ REAL rxy_d, rbx_d, rby_d, rscale_d;
__constant__ REAL invdx_d, invdy_d;
__constant__ REAL tc_d, alpha_d, beta_d, delta_d, eps_d, mu_d, gamma_d, theta_d;
__constant__ REAL boundaryVal_d;
__constant__ bool solidSwitch_d, neumannBC_d, gateDiff_d, anisotropy_d, tipGrad_d;
__constant__ int tipOffsetX_d, tipOffsetY_d;
__constant__ float minVarColor_d, maxVarColor_d;
__constant__ float tipx0_d, tipy0_d;
__constant__ REAL Uth_d, conTh1_d, conTh2_d, conTh3_d;
__constant__ int lap4_d, timeIntOrder_d;

// helper_functions.cu:164-178 calls SOIL (screenshots); never reached headless.
extern "C" int SOIL_save_screenshot(const char *, int, int, int, int, int) { return 0; }

static int g_err = 0;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "ref_harness: %s -> %s\n", #x, cudaGetErrorString(e_)); g_err = 1; } } while (0)
#define SYM(sym, val) CK(cudaMemcpyToSymbol(sym, &(val), sizeof(val)))

static dim3 grid0D, block0D, grid1D, block1D, grid2D, block2D;

extern "C" {

// Fill `param` with the reference's own parameterSetup() (saveFiles.cu:105-231), override
// with the scalars of *p, upload every constant exactly as main.cu:309-402.
int yref_init(const yh_params *p) {
  g_err = 0;
  // tip_wrapper / countour_wrapper read an UNINITIALISED malloc'd int as a cudaMemset size
  // (tipTracker.cu:574-577, defect B4).  M_PERTURB = 0xff makes glibc hand out zero-filled
  // chunks, so that memset is a deterministic no-op instead of an out-of-bounds write.
  mallopt(M_PERTURB, 0xff);
  param = parameterSetup(param);
  param.nx = p->nx; param.ny = p->ny;
  param.solidSwitch = p->solidSwitch; param.neumannBC = p->neumannBC;
  param.gateDiff = p->gateDiff; param.anisotropy = p->anisotropy; param.lap4 = p->lap4;
  param.timeIntOrder = p->timeIntOrder; param.tipGrad = p->tipGrad;
  param.tipAlgorithm = p->tipAlgorithm;
  param.tipOffsetX = p->tipOffsetX; param.tipOffsetY = p->tipOffsetY;
  param.tipx = p->tipx0; param.tipy = p->tipy0;
  param.dt = p->dt; param.hx = p->hx; param.hy = p->hy; param.Lx = p->Lx; param.Ly = p->Ly;
  param.rx = p->rx; param.ry = p->ry; param.rxy = p->rxy; param.rbx = p->rbx; param.rby = p->rby;
  param.rscale = p->rscale; param.qx4 = p->qx4; param.qy4 = p->qy4; param.fx4 = p->fx4;
  param.fy4 = p->fy4; param.invdx = p->invdx; param.invdy = p->invdy;
  param.tc = p->tc; param.alpha = p->alpha; param.beta = p->beta; param.gamma = p->gamma;
  param.delta = p->delta; param.eps = p->eps; param.mu = p->mu; param.theta = p->theta;
  param.boundaryVal = p->boundaryVal; param.Uth = p->Uth;
  param.point = make_int2(param.nx / 2, param.ny / 2);
  param.wnx = param.nx; param.wny = param.ny;
  param.memSize = (int)((size_t)param.nx * param.ny * sizeof(REAL));   // overflows at 16384^2 (F12); unused here

  grid0D = dim3(1, 1, 1); block0D = dim3(1, 1, 1);                    // main.cu:167-172
  grid1D = dim3(GRIDSIZE_1D, 1, 1); block1D = dim3(BLOCKSIZE_1D, 1, 1);
  grid2D = dim3(iDivUp(param.nx, BLOCK_DIM_X), iDivUp(param.ny, BLOCK_DIM_Y), 1);
  block2D = dim3(BLOCK_DIM_X, BLOCK_DIM_Y, 1);

  SYM(nx_d, param.nx); SYM(ny_d, param.ny);
  SYM(rx_d, param.rx); SYM(ry_d, param.ry); SYM(qx4_d, param.qx4); SYM(qy4_d, param.qy4);
  SYM(fx4_d, param.fx4); SYM(fy4_d, param.fy4); SYM(hx_d, param.hx); SYM(hy_d, param.hy);
  SYM(dt_d, param.dt); SYM(invdx_d, param.invdx); SYM(invdy_d, param.invdy);
  SYM(Lx_d, param.Lx); SYM(Ly_d, param.Ly); SYM(rxy_d, param.rxy); SYM(rbx_d, param.rbx);
  SYM(rby_d, param.rby); SYM(rscale_d, param.rscale); SYM(boundaryVal_d, param.boundaryVal);
  SYM(solidSwitch_d, param.solidSwitch); SYM(neumannBC_d, param.neumannBC);
  SYM(gateDiff_d, param.gateDiff); SYM(anisotropy_d, param.anisotropy);
  SYM(tipGrad_d, param.tipGrad); SYM(lap4_d, param.lap4); SYM(timeIntOrder_d, param.timeIntOrder);
  SYM(tipOffsetX_d, param.tipOffsetX); SYM(tipOffsetY_d, param.tipOffsetY);
  SYM(minVarColor_d, param.minVarColor); SYM(maxVarColor_d, param.maxVarColor);
  SYM(tipx0_d, param.tipx); SYM(tipy0_d, param.tipy);
  SYM(conTh1_d, param.contourThresh1); SYM(conTh2_d, param.contourThresh2);
  SYM(conTh3_d, param.contourThresh3); SYM(Uth_d, param.Uth);
  SYM(tc_d, param.tc); SYM(alpha_d, param.alpha); SYM(beta_d, param.beta);
  SYM(gamma_d, param.gamma); SYM(delta_d, param.delta); SYM(eps_d, param.eps);
  SYM(mu_d, param.mu); SYM(theta_d, param.theta);
  CK(cudaDeviceSynchronize());
  return g_err ? -2 : 0;
}

static REAL *dalloc(size_t n, const REAL *h) {
  REAL *d = nullptr;
  CK(cudaMalloc(&d, n * sizeof(REAL)));
  if (h) CK(cudaMemcpy(d, h, n * sizeof(REAL), cudaMemcpyHostToDevice));
  else CK(cudaMemset(d, 0, n * sizeof(REAL)));
  return d;
}
static bool *balloc(size_t n, const uint8_t *h) {
  bool *d = nullptr;
  CK(cudaMalloc(&d, n * sizeof(bool)));
  if (h) CK(cudaMemcpy(d, h, n, cudaMemcpyHostToDevice));
  else CK(cudaMemset(d, 0, n));
  return d;
}

// N x { reactionDiffusion_wrapper ; swapSoA }  (main.cu:879-882).  u,v updated in place on
// the host; velTan (optional) receives the last step's velTan.  mode 1 adds the as-shipped
// per-step singleCell_wrapper (blocking 16-byte D2H, main.cu:1040).  Returns elapsed ms of
// the step loop (CUDA events) or <0 on error.
float yref_rd_run(double *u_h, double *v_h, double *vtu_h, double *vtv_h,
                  const uint8_t *solid_h, int nsteps, int stim_mouse, int px, int py, int mode,
                  int copy_back) {
  g_err = 0;
  const size_t n = (size_t)param.nx * param.ny;
  stateVar gin, gout, J, vt;
  gin.u = dalloc(n, u_h); gin.v = dalloc(n, v_h);
  gout.u = dalloc(n, nullptr); gout.v = dalloc(n, nullptr);
  J.u = dalloc(n, nullptr); J.v = dalloc(n, nullptr);
  vt.u = dalloc(n, nullptr); vt.v = dalloc(n, nullptr);
  bool *solid_d = balloc(n, solid_h);
  REAL *stim_d = dalloc(n, nullptr);
  REAL *pt_d = dalloc(2, nullptr);
  REAL pt_h[2];
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int s = 0; s < nsteps; s++) {
    reactionDiffusion_wrapper(pitch, grid2D, block2D, gout, gin, J, vt, false, solid_d, false,
                              stim_d, stim_mouse != 0, make_int2(px, py));
    swapSoA(&gin, &gout);
    if (mode == 1) singleCell_wrapper(pitch, grid0D, block0D, gout, 2, pt_h, pt_d, make_int2(px, py));
  }
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGetLastError());
  if (copy_back) {
    CK(cudaMemcpy(u_h, gin.u, n * sizeof(REAL), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(v_h, gin.v, n * sizeof(REAL), cudaMemcpyDeviceToHost));
    if (vtu_h) CK(cudaMemcpy(vtu_h, vt.u, n * sizeof(REAL), cudaMemcpyDeviceToHost));
    if (vtv_h) CK(cudaMemcpy(vtv_h, vt.v, n * sizeof(REAL), cudaMemcpyDeviceToHost));
  }
  cudaFree(gin.u); cudaFree(gin.v); cudaFree(gout.u); cudaFree(gout.v);
  cudaFree(J.u); cudaFree(J.v); cudaFree(vt.u); cudaFree(vt.v);
  cudaFree(solid_d); cudaFree(stim_d); cudaFree(pt_d);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return g_err ? -1.f : ms;
}

// tip_wrapper (tipTracker.cu:569-611).  Caller convention of main.cu:963: first stateVar is
// the PRESENT field, second the PAST field.  Returns the tip count (unsorted list) or <0.
int yref_tip(const double *u_present_h, const double *u_past_h, double t, int algorithm,
             yh_tip *tips_h, int capacity, uint8_t *tip_plot_h) {
  g_err = 0;
  const size_t n = (size_t)param.nx * param.ny;
  stateVar pres, past, vt;
  pres.u = dalloc(n, u_present_h); pres.v = nullptr;
  past.u = dalloc(n, u_past_h); past.v = nullptr;
  vt.u = vt.v = nullptr;
  bool *plot_d = balloc(n, nullptr);
  int *cnt_d = nullptr; vec5dyn *vec_d = nullptr;
  CK(cudaMalloc(&cnt_d, sizeof(int)));
  CK(cudaMalloc(&vec_d, sizeof(vec5dyn) * (size_t)TIPVECSIZE));
  CK(cudaMemset(vec_d, 0, sizeof(vec5dyn) * (size_t)TIPVECSIZE));
  tip_wrapper(pitch, grid2D, block2D, pres, past, vt, t, algorithm, false, plot_d, cnt_d, vec_d);
  CK(cudaDeviceSynchronize());
  cudaGetLastError();   // the wrapper's cudaMemset of an uninitialised size may fail (B4); clear it
  int cnt = 0;
  CK(cudaMemcpy(&cnt, cnt_d, sizeof(int), cudaMemcpyDeviceToHost));
  int m = cnt < capacity ? cnt : capacity;
  if (m > 0) CK(cudaMemcpy(tips_h, vec_d, sizeof(vec5dyn) * (size_t)m, cudaMemcpyDeviceToHost));
  if (tip_plot_h) CK(cudaMemcpy(tip_plot_h, plot_d, n, cudaMemcpyDeviceToHost));
  cudaFree(pres.u); cudaFree(past.u); cudaFree(plot_d); cudaFree(cnt_d); cudaFree(vec_d);
  return g_err ? -1 : cnt;
}

// slice_wrapper (scheme 2, reduceSymStart = true) followed by trapz_wrapper
// (main.cu:906,923).  The disc centre comes from a one-entry tip list (count != 0) or from
// tipx0/tipy0 (count == 0).  slices_h (optional): 12 arrays, slice then slice0.
int yref_slice_trapz(const double *u_h, const double *v_h, const double *advx_h,
                     const double *advy_h, const double *vtu_h, const double *vtv_h,
                     float tipx, float tipy, int count, double *integrals_h, double *slices_h) {
  g_err = 0;
  const size_t n = (size_t)param.nx * param.ny;
  stateVar g, vt; advVar adv; sliceVar s, s0;
  g.u = dalloc(n, u_h); g.v = dalloc(n, v_h);
  vt.u = dalloc(n, vtu_h); vt.v = dalloc(n, vtv_h);
  adv.x = dalloc(n, advx_h); adv.y = dalloc(n, advy_h);
  REAL **sp[12] = {&s.ux, &s.uy, &s.ut, &s.vx, &s.vy, &s.vt, &s0.ux, &s0.uy, &s0.ut, &s0.vx, &s0.vy, &s0.vt};
  for (int k = 0; k < 12; k++) *sp[k] = dalloc(n, nullptr);
  bool *area_d = balloc(n, nullptr);
  REAL *coeff_d = dalloc(n, nullptr);
  int *cnt_d = nullptr; vec5dyn *vec_d = nullptr;
  CK(cudaMalloc(&cnt_d, sizeof(int)));
  CK(cudaMalloc(&vec_d, sizeof(vec5dyn) * 4));
  int one = 1; vec5dyn tv = {tipx, tipy, 0.f, 0.f, 0.f};
  CK(cudaMemcpy(cnt_d, &one, sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(vec_d, &tv, sizeof(vec5dyn), cudaMemcpyHostToDevice));
  slice_wrapper(pitch, grid2D, block2D, g, s, s0, true, true, adv, 2, area_d, cnt_d, vec_d, count);
  trapz_wrapper(grid1D, block1D, s, s0, vt, integrals_h, coeff_d, cnt_d, vec_d, count);
  CK(cudaDeviceSynchronize());
  if (slices_h)
    for (int k = 0; k < 12; k++)
      CK(cudaMemcpy(slices_h + (size_t)k * n, *sp[k], n * sizeof(REAL), cudaMemcpyDeviceToHost));
  cudaFree(g.u); cudaFree(g.v); cudaFree(vt.u); cudaFree(vt.v); cudaFree(adv.x); cudaFree(adv.y);
  for (int k = 0; k < 12; k++) cudaFree(*sp[k]);
  cudaFree(area_d); cudaFree(coeff_d); cudaFree(cnt_d); cudaFree(vec_d);
  return g_err ? -1 : 0;
}

int yref_solve_matrix(const double c_in[3], const double phi[3], double *Int, double c_out[3]) {
  REAL3 c = {c_in[0], c_in[1], c_in[2]}, ph = {phi[0], phi[1], phi[2]};
  REAL3 r = solve_matrix(c, ph, Int);
  c_out[0] = r.x; c_out[1] = r.y; c_out[2] = r.t;
  return 0;
}

int yref_cxy(const double c[3], const double phi[3], const uint8_t *solid_h,
             double *advx_h, double *advy_h) {
  g_err = 0;
  const size_t n = (size_t)param.nx * param.ny;
  advVar adv; adv.x = dalloc(n, nullptr); adv.y = dalloc(n, nullptr);
  bool *solid_d = balloc(n, solid_h);
  REAL3 cc = {c[0], c[1], c[2]}, ph = {phi[0], phi[1], phi[2]};
  Cxy_field_wrapper(pitch, grid2D, block2D, adv, cc, ph, solid_d);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(advx_h, adv.x, n * sizeof(REAL), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(advy_h, adv.y, n * sizeof(REAL), cudaMemcpyDeviceToHost));
  cudaFree(adv.x); cudaFree(adv.y); cudaFree(solid_d);
  return g_err ? -1 : 0;
}

// advFDBFECC_wrapper (advFDBFECC.cu:355-362); racy in the reference (B2).
// `repeats` > 1 calls the wrapper again on the SAME buffers: uf is race-free (a pure function
// of g_in), so the second call reads correct uf everywhere and the third correct ue -- after 3
// identical calls the racy kernel has converged to the synchronous-sweep result.
int yref_bfecc(const double *u_h, const double *v_h, const double *advx_h, const double *advy_h,
               const uint8_t *solid_h, double *uo_h, double *vo_h, int repeats) {
  g_err = 0;
  const size_t n = (size_t)param.nx * param.ny;
  stateVar gi, go, uf, ub, ue; advVar adv;
  gi.u = dalloc(n, u_h); gi.v = dalloc(n, v_h);
  go.u = dalloc(n, nullptr); go.v = dalloc(n, nullptr);
  uf.u = dalloc(n, nullptr); uf.v = dalloc(n, nullptr);
  ub.u = dalloc(n, nullptr); ub.v = dalloc(n, nullptr);
  ue.u = dalloc(n, nullptr); ue.v = dalloc(n, nullptr);
  adv.x = dalloc(n, advx_h); adv.y = dalloc(n, advy_h);
  bool *solid_d = balloc(n, solid_h);
  for (int r = 0; r < (repeats > 0 ? repeats : 1); r++) {
    advFDBFECC_wrapper(pitch, grid2D, block2D, go, gi, adv, uf, ub, ue, solid_d);
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(uo_h, go.u, n * sizeof(REAL), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(vo_h, go.v, n * sizeof(REAL), cudaMemcpyDeviceToHost));
  REAL *all[] = {gi.u, gi.v, go.u, go.v, uf.u, uf.v, ub.u, ub.v, ue.u, ue.v, adv.x, adv.y};
  for (REAL *q : all) cudaFree(q);
  cudaFree(solid_d);
  return g_err ? -1 : 0;
}

// sAPD_wrapper (spaceAPD.cu:376-384) over a sequence of nframes fields u_seq[k] (host),
// called with (uold=u_seq[k], unew=u_seq[k+1], count=count0+k).  State arrays are zeroed
// first (B12).  Outputs: APD1, APD2, sAPD, dAPD, back, front (6*n doubles) and first (n bytes).
int yref_sapd(const double *u_seq_h, int nframes, int count0, const uint8_t *stimArea_h,
              int stimulate, double *out6_h, uint8_t *first_h) {
  g_err = 0;
  const size_t n = (size_t)param.nx * param.ny;
  REAL *seq = dalloc(n * nframes, u_seq_h);
  REAL *st[6]; for (int k = 0; k < 6; k++) st[k] = dalloc(n, nullptr);
  bool *first_d = balloc(n, nullptr), *area_d = balloc(n, stimArea_h);
  for (int k = 0; k + 1 < nframes; k++)
    sAPD_wrapper(pitch, grid1D, block1D, count0 + k, seq + (size_t)k * n, seq + (size_t)(k + 1) * n,
                 st[0], st[1], st[2], st[3], st[4], st[5], first_d, area_d, stimulate != 0);
  CK(cudaDeviceSynchronize());
  for (int k = 0; k < 6; k++)
    CK(cudaMemcpy(out6_h + (size_t)k * n, st[k], n * sizeof(REAL), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(first_h, first_d, n, cudaMemcpyDeviceToHost));
  cudaFree(seq); for (int k = 0; k < 6; k++) cudaFree(st[k]);
  cudaFree(first_d); cudaFree(area_d);
  return g_err ? -1 : 0;
}

// display() with param.reduceSym (main.cu:894-954), the reference's own wrappers and host solve,
// exactly as shipped: 12 blocking D2H + cudaMalloc/cudaFree per trapz_wrapper call, sticky
// reduceSymStart (B6), the extra solve/Cxy/slice at count == 0.  c_phi_h: 6 doubles per step
// (clist/philist).  Returns elapsed ms of the loop (CUDA events) or <0.
float yref_sr_run(double *u_h, double *v_h, int nsteps, double *c_phi_h, int copy_back) {
  g_err = 0;
  const size_t n = (size_t)param.nx * param.ny;
  stateVar gateIn_d, gateOut_d, J_d, velTan, uf, ub, ue; advVar advect; sliceVar sl, sl0;
  gateIn_d.u = dalloc(n, u_h); gateIn_d.v = dalloc(n, v_h);
  gateOut_d.u = dalloc(n, nullptr); gateOut_d.v = dalloc(n, nullptr);
  J_d.u = dalloc(n, nullptr); J_d.v = dalloc(n, nullptr);
  velTan.u = dalloc(n, nullptr); velTan.v = dalloc(n, nullptr);
  uf.u = dalloc(n, nullptr); uf.v = dalloc(n, nullptr); ub.u = dalloc(n, nullptr); ub.v = dalloc(n, nullptr);
  ue.u = dalloc(n, nullptr); ue.v = dalloc(n, nullptr);
  advect.x = dalloc(n, nullptr); advect.y = dalloc(n, nullptr);
  REAL **sp[12] = {&sl.ux, &sl.uy, &sl.ut, &sl.vx, &sl.vy, &sl.vt, &sl0.ux, &sl0.uy, &sl0.ut, &sl0.vx, &sl0.vy, &sl0.vt};
  for (int k = 0; k < 12; k++) *sp[k] = dalloc(n, nullptr);
  bool *solid_d = balloc(n, nullptr), *tip_plot = balloc(n, nullptr), *area_d = balloc(n, nullptr);
  CK(cudaMemset(solid_d, 1, n));
  REAL *stim_d = dalloc(n, nullptr), *coeff_d = dalloc(n, nullptr);
  int *tip_count_d = nullptr; vec5dyn *tip_vector_d = nullptr;
  CK(cudaMalloc(&tip_count_d, sizeof(int))); CK(cudaMemset(tip_count_d, 0, sizeof(int)));
  CK(cudaMalloc(&tip_vector_d, sizeof(vec5dyn) * (size_t)TIPVECSIZE));
  CK(cudaMemset(tip_vector_d, 0, sizeof(vec5dyn) * (size_t)TIPVECSIZE));
  REAL integrals[12];
  REAL3 c = {0, 0, 0}, phi = {0, 0, 0};
  param.reduceSym = true; param.reduceSymStart = true; param.count = 0; param.physicalTime = 0.0;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int it = 0; it < nsteps; it++) {
    reactionDiffusion_wrapper(pitch, grid2D, block2D, gateOut_d, gateIn_d, J_d, velTan, param.reduceSym, solid_d,
                              false, stim_d, false, param.point);
    tip_wrapper(pitch, grid2D, block2D, gateIn_d, gateOut_d, velTan, param.physicalTime, param.tipAlgorithm,
                param.recordTip, tip_plot, tip_count_d, tip_vector_d);
    if (c_phi_h) { REAL *o = c_phi_h + 6 * it; o[0] = c.x; o[1] = c.y; o[2] = c.t; o[3] = phi.x; o[4] = phi.y; o[5] = phi.t; }
    slice_wrapper(pitch, grid2D, block2D, gateIn_d, sl, sl0, param.reduceSym, param.reduceSymStart, advect, 2,
                  area_d, tip_count_d, tip_vector_d, param.count);
    if (param.count == 0) {
      trapz_wrapper(grid1D, block1D, sl, sl0, velTan, integrals, coeff_d, tip_count_d, tip_vector_d, param.count);
      c = solve_matrix(c, phi, integrals);
      Cxy_field_wrapper(pitch, grid2D, block2D, advect, c, phi, solid_d);
      slice_wrapper(pitch, grid2D, block2D, gateIn_d, sl, sl0, param.reduceSym, param.reduceSymStart, advect, 2,
                    area_d, tip_count_d, tip_vector_d, param.count);
    }
    trapz_wrapper(grid1D, block1D, sl, sl0, velTan, integrals, coeff_d, tip_count_d, tip_vector_d, param.count);
    c = solve_matrix(c, phi, integrals);
    Cxy_field_wrapper(pitch, grid2D, block2D, advect, c, phi, solid_d);
    advFDBFECC_wrapper(pitch, grid2D, block2D, gateIn_d, gateOut_d, advect, uf, ub, ue, solid_d);
    phi.x = phi.x + c.x * param.dt; phi.y = phi.y + c.y * param.dt; phi.t = phi.t + c.t * param.dt;
    param.count++;
    param.physicalTime = param.dt * (REAL)param.count;
  }
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaGetLastError();
  if (copy_back) {
    CK(cudaMemcpy(u_h, gateIn_d.u, n * sizeof(REAL), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(v_h, gateIn_d.v, n * sizeof(REAL), cudaMemcpyDeviceToHost));
  }
  REAL *all[] = {gateIn_d.u, gateIn_d.v, gateOut_d.u, gateOut_d.v, J_d.u, J_d.v, velTan.u, velTan.v, uf.u, uf.v,
                 ub.u, ub.v, ue.u, ue.v, advect.x, advect.y, stim_d, coeff_d};
  for (REAL *q : all) cudaFree(q);
  for (int k = 0; k < 12; k++) cudaFree(*sp[k]);
  cudaFree(solid_d); cudaFree(tip_plot); cudaFree(area_d); cudaFree(tip_count_d); cudaFree(tip_vector_d);
  param.reduceSym = false; param.reduceSymStart = false;
  return g_err ? -1.f : ms;
}

// contourMode == 1 loop for ONE sheet (main.cu:879-885, 1035, 1040): RD with the paced disc
// stimulus, swap, sAPD_wrapper(count, gateIn.u, gateOut.u, ..., stimulate = true), and (mode 1)
// the per-step singleCell_wrapper.  apd_h: APD1 then APD2 (2*n doubles).
float yref_apd_run(double *u_h, double *v_h, int nsteps, int period_it, int duration_it,
                   const uint8_t *stimArea_h, double *apd_h, int mode) {
  g_err = 0;
  const size_t n = (size_t)param.nx * param.ny;
  stateVar gin, gout, J, vt;
  gin.u = dalloc(n, u_h); gin.v = dalloc(n, v_h); gout.u = dalloc(n, nullptr); gout.v = dalloc(n, nullptr);
  J.u = dalloc(n, nullptr); J.v = dalloc(n, nullptr); vt.u = dalloc(n, nullptr); vt.v = dalloc(n, nullptr);
  REAL *st[6]; for (int k = 0; k < 6; k++) st[k] = dalloc(n, nullptr);
  bool *first_d = balloc(n, nullptr), *area_d = balloc(n, stimArea_h), *solid_d = balloc(n, nullptr);
  REAL *stim_d = dalloc(n, nullptr), *pt_d = dalloc(2, nullptr);
  REAL pt_h[2];
  int count = 0;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int s = 0; s < nsteps; s++) {
    const bool on = period_it > 0 && (count % period_it) <= duration_it;
    reactionDiffusion_wrapper(pitch, grid2D, block2D, gout, gin, J, vt, false, solid_d, false, stim_d, on, param.point);
    swapSoA(&gin, &gout);
    count++;
    sAPD_wrapper(pitch, grid1D, block1D, count, gin.u, gout.u, st[0], st[1], st[2], st[3], st[4], st[5], first_d,
                 area_d, true);
    if (mode == 1) singleCell_wrapper(pitch, grid0D, block0D, gout, 2, pt_h, pt_d, param.point);
  }
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaMemcpy(u_h, gin.u, n * sizeof(REAL), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(v_h, gin.v, n * sizeof(REAL), cudaMemcpyDeviceToHost));
  if (apd_h) {
    CK(cudaMemcpy(apd_h, st[0], n * sizeof(REAL), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(apd_h + n, st[1], n * sizeof(REAL), cudaMemcpyDeviceToHost));
  }
  REAL *all[] = {gin.u, gin.v, gout.u, gout.v, J.u, J.v, vt.u, vt.v, stim_d, pt_d};
  for (REAL *q : all) cudaFree(q);
  for (int k = 0; k < 6; k++) cudaFree(st[k]);
  cudaFree(first_d); cudaFree(area_d); cudaFree(solid_d);
  return g_err ? -1.f : ms;
}

// countour_wrapper (spaceAPD.cu:256-276).  field2 is allocated with nx+1 zero cells of padding
// because the kernel reads I2D(i+1,j) / I2D(i,j+1) past the end of the array on the last row
// (:52-53); fixtures keep the last row free of hits, where the reference's result is undefined.
// Returns the point count (list in atomicAdd order) or <0.
int yref_contour(const double *field1_h, const double *field2_h, const uint8_t *stimArea_h,
                 float t, int mode, yh_contour_pt *pts_h, int capacity, uint8_t *plot_h) {
  g_err = 0;
  const size_t n = (size_t)param.nx * param.ny;
  REAL *f1 = dalloc(n + param.nx + 1, nullptr), *f2 = dalloc(n + param.nx + 1, nullptr);
  if (field1_h) CK(cudaMemcpy(f1, field1_h, n * sizeof(REAL), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(f2, field2_h, n * sizeof(REAL), cudaMemcpyHostToDevice));
  bool *plot_d = balloc(n, nullptr), *area_d = balloc(n, stimArea_h);
  int *cnt_d = nullptr; float3 *vec_d = nullptr;
  CK(cudaMalloc(&cnt_d, sizeof(int)));
  CK(cudaMalloc(&vec_d, sizeof(float3) * 2 * n));
  CK(cudaMemset(vec_d, 0, sizeof(float3) * 2 * n));
  countour_wrapper(pitch, grid2D, block2D, f1, f2, plot_d, area_d, cnt_d, vec_d, t, mode);
  CK(cudaDeviceSynchronize());
  cudaGetLastError();   // cudaMemset of an uninitialised size (spaceAPD.cu:262-265, defect B4)
  int cnt = 0;
  CK(cudaMemcpy(&cnt, cnt_d, sizeof(int), cudaMemcpyDeviceToHost));
  int m = cnt < capacity ? cnt : capacity;
  if (m > 0) CK(cudaMemcpy(pts_h, vec_d, sizeof(float3) * (size_t)m, cudaMemcpyDeviceToHost));
  if (plot_h) CK(cudaMemcpy(plot_h, plot_d, n, cudaMemcpyDeviceToHost));
  cudaFree(f1); cudaFree(f2); cudaFree(plot_d); cudaFree(area_d); cudaFree(cnt_d); cudaFree(vec_d);
  return g_err ? -1 : cnt;
}

}  // extern "C"
