// yh_common.cuh -- shared device/host definitions of libyolohtli_b200 (sm_100a only).
//
// Parameters travel to every kernel BY VALUE in one __grid_constant__ block (YhK) instead
// of the reference's ~45 __constant__ symbols resolved at device-link time
// (main.cu:40-51, 309-402), so no -rdc and no global state: two simulations with different
// parameters can be in flight on different streams.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/yolohtli_abi.h"

// Kernel-side copy of yh_params (POD, passed by value).
struct YhK {
  int nx, ny, nyg, jg0;
  int solidSwitch, neumannBC, gateDiff, anisotropy, lap4, timeIntOrder, tipGrad;
  int tipOffX, tipOffY;
  int stim, px, py;          // live disc stimulus (reactionDiffusion.cu:54-61)
  int row0, row1;            // local rows to write
  double dt, hx, hy, rx, ry, rxy, rbx, rby, rscale, qx4, qy4, fx4, fy4, invdx, invdy;
  double tc, alpha, beta, gamma, delta, eps, mu, theta, boundaryVal, Uth;
};

static inline YhK yh_make_k(const yh_params *p) {
  YhK k;
  k.nx = p->nx; k.ny = p->ny; k.nyg = p->ny_global; k.jg0 = p->jg0;
  k.solidSwitch = p->solidSwitch; k.neumannBC = p->neumannBC; k.gateDiff = p->gateDiff;
  k.anisotropy = p->anisotropy; k.lap4 = p->lap4; k.timeIntOrder = p->timeIntOrder;
  k.tipGrad = p->tipGrad; k.tipOffX = p->tipOffsetX; k.tipOffY = p->tipOffsetY;
  k.stim = 0; k.px = 0; k.py = 0; k.row0 = 0; k.row1 = p->ny;
  k.dt = p->dt; k.hx = p->hx; k.hy = p->hy; k.rx = p->rx; k.ry = p->ry; k.rxy = p->rxy;
  k.rbx = p->rbx; k.rby = p->rby; k.rscale = p->rscale; k.qx4 = p->qx4; k.qy4 = p->qy4;
  k.fx4 = p->fx4; k.fy4 = p->fy4; k.invdx = p->invdx; k.invdy = p->invdy;
  k.tc = p->tc; k.alpha = p->alpha; k.beta = p->beta; k.gamma = p->gamma; k.delta = p->delta;
  k.eps = p->eps; k.mu = p->mu; k.theta = p->theta; k.boundaryVal = p->boundaryVal;
  k.Uth = p->Uth;
  return k;
}

// ---- error plumbing (abi.cu) -----------------------------------------------------------
void yh_set_error(const char *fmt, ...);
int yh_check_device(void);   // YH_OK or YH_ERR_NO_DEVICE

#define YH_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      yh_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return YH_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

#define YH_REQUIRE(cond, msg)                                  \
  do {                                                         \
    if (!(cond)) {                                             \
      yh_set_error("%s: %s", __func__, msg);                   \
      return YH_ERR_INVALID_ARG;                               \
    }                                                          \
  } while (0)

#define YH_LAUNCH_CHECK() YH_CUDA(cudaGetLastError())

// Kernel launch of the RD step path, or -- in PRELOAD mode -- only the load of the kernel's code.
// CUDA loads a kernel lazily at its first launch and that load can wait until the device is idle.  The
// slab driver keeps a kernel resident that waits for a neighbour slab (csrc/slab.cu); a first launch of
// some RD kernel variant issued by the same host thread meanwhile would block behind it while the
// neighbour's work is not enqueued yet.  So the driver walks its launch sequence once with
// yh_preload_only set (no launches, cudaFuncGetAttributes loads the code) before it enqueues anything.
extern thread_local int yh_preload_only;
#define YH_LAUNCH(kfn, grd, blk, smem, st, ...)                    \
  do {                                                             \
    if (yh_preload_only) {                                         \
      cudaFuncAttributes fa__;                                     \
      YH_CUDA(cudaFuncGetAttributes(&fa__, kfn));                  \
    } else {                                                       \
      kfn<<<grd, blk, smem, st>>>(__VA_ARGS__);                    \
      YH_LAUNCH_CHECK();                                           \
    }                                                              \
  } while (0)

// internal workspace (per device, grown on demand, abi.cu)
int yh_workspace(size_t bytes, void **ptr, int slot);
unsigned long long yh_workspace_generation(void);
unsigned yh_next_epoch(void);   // launch epoch of the ordered-compaction kernels, process-wide

// nsteps x {RD step; swapSoA} on (uA,vA) <-> (uB,vB), CUDA-graph replay for small whole sheets
// (abi.cu).  *last_T (optional) = time steps of the final pass.
int yh_advance_whole(const yh_params *p, const YhK &k, int nsteps, int tb, int canon_in, double *uA,
                     double *vA, double *uB, double *vB, const uint8_t *solid, const uint8_t *pat,
                     int row0, int row1, int *result_in_B, int *last_T, cudaStream_t st);

// solve_matrix, symmetryReduction.cu:386-416: rotate the first two columns by phi.t, then the 3x3
// elimination without pivoting, operation for operation.  Host (libm cos/sin, as the reference)
// and device (libdevice cos/sin; yh_sim_run_sr_device) share the text.
__host__ __device__ inline void yh_solve3(const double *Int, double cs, double sn, double *c_out) {
  double a1, a2, a3, b1, b2, b3, C1, C2, C3, d1, d2, d3;
  double b2p, b3p, c2p, c3p, c3pp, d2p, d3p, d3pp, x1, x2, x3;
  a1 = Int[0] * cs + Int[1] * sn; a2 = Int[1] * cs - Int[0] * sn; a3 = Int[2];
  b1 = Int[3] * cs + Int[4] * sn; b2 = Int[4] * cs - Int[3] * sn; b3 = Int[5];
  C1 = Int[6] * cs + Int[7] * sn; C2 = Int[7] * cs - Int[6] * sn; C3 = Int[8];
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
}

// Device-resident symmetry-reduction state (yh_sim_run_sr_device): c[3], phi[3], cos/sin(phi.t)
#define YH_SR_C 0
#define YH_SR_PHI 3
#define YH_SR_CS 6
#define YH_SR_SN 7
#define YH_SR_STEP 8     /* param.count as a double: advanced by the closing kernel of every step */
#define YH_SR_WORDS 10
// log_base / step0: the (c, phi) record of step n goes to log_base + 6*(n - step0); the deferred
// phi += c*dt is applied for n > step0.  Every argument is the same for every step of a run, so
// the step can be replayed from a CUDA graph.
int yh_sr_integrals_solve_device(const yh_params *p, const double *u, const double *v,
                                 const double *vtu, const double *vtv, const double *ax,
                                 const double *ay, const int *tip_count, const yh_tip *tv,
                                 double *sr_state, double *log_base, double step0, cudaStream_t st);
// tip_wrapper with the time tag and the look-back epoch taken from the device-resident step counter
int yh_tip_track_device_step(const yh_params *p, const double *u_past, const double *u_present,
                             int *tip_count, yh_tip *tip_vector, int capacity, const double *sr_state,
                             int steps_ahead, cudaStream_t st);
int yh_sr_flush_phi_device(double *sr_state, double dt_phi, cudaStream_t st);
int yh_advect_bfecc_device_c(const yh_params *p, const double *u_in, const double *v_in, double *u_out,
                             double *v_out, const double *sr_state, double *adv_x, double *adv_y,
                             const uint8_t *solid, cudaStream_t st);

int yh_graphs_enabled(long long cells);
int yh_arithmetic(void);   // abi.cu: YH_ARITH_EXACT | YH_ARITH_FAST (yh_set_arithmetic, or YH_ARITH = exact | fast)   // abi.cu: YH_GRAPHS = 0 | 1 override, else small sheets only

// ---- device helpers --------------------------------------------------------------------
// Neumann mirror index (the rule of coord_i/coord_j, helper_functions.cu:69-79).
__device__ __forceinline__ int yh_mir(int i, int n) {
  return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i);
}

// The ionic model (reactionDiffusion.cu:131-141); scs = inside the live stimulus disc.
__device__ __forceinline__ double yh_Isum(const YhK &k, double u, double v, bool scs) {
  return -(k.mu * u * (1.0 - u) * (u - k.alpha) - u * v) - (scs ? 24.7 : 0.0);
}
__device__ __forceinline__ double yh_Iv(const YhK &k, double u, double v) {
  return -(k.eps * (k.delta * (u - k.gamma) * (k.beta - u) - v - k.theta));
}
// i, j are GLOBAL cell coordinates.
__device__ __forceinline__ bool yh_scs(const YhK &k, int i, int j) {
  if (!k.stim) return false;
  int ic = i - k.nx / 2, jc = j - k.nyg / 2;
  int cx = k.px - k.nx / 2, cy = k.py - k.nyg / 2;
  return ((ic - cx) * (ic - cx) + (jc - cy) * (jc - cy)) < 400;
}

// disc test only (caller already knows the stimulus is on)
__device__ __forceinline__ bool yh_scs_on(const YhK &k, int i, int j) {
  int ic = i - k.nx / 2, jc = j - k.nyg / 2;
  int cx = k.px - k.nx / 2, cy = k.py - k.nyg / 2;
  return ((ic - cx) * (ic - cx) + (jc - cy) * (jc - cy)) < 400;
}

// APD bookkeeping state (sAPD_kernel, spaceAPD.cu:278-374) for the fused epilogue of the fast
// kernel; all pointers NULL = off.  Arrays span the whole batch (sheet z at offset z*stride).
struct YhApd {
  double *APD1, *APD2, *sAPD, *dAPD, *back, *front;
  uint8_t *first;
  const uint8_t *stimArea;
  int stimulate;
};

// Rows a stage reads beyond the rows it writes.  1, except for the anisotropic no-flux corner corrections,
// which read rows j +- 2 at the columns x = 0 and x = nx-1 (reactionDiffusion.cu:290-304).
static inline int yh_rd_radius(const yh_params *p) {
  return (p->anisotropy && p->neumannBC && !p->solidSwitch) ? 2 : 1;
}

// ---- kernel launchers (one per .cu) ------------------------------------------------------
int yh_launch_rd_generic(const YhK &k, const double *u_in, const double *v_in, double *u_out,
                         double *v_out, double *vtu, double *vtv, const uint8_t *solid,
                         cudaStream_t st);
// Temporally blocked Euler/5-point path; returns YH_ERR_UNSUPPORTED when the mode is not
// covered so the caller can fall back to the generic kernels.
int yh_rd_fast_supported(const YhK &k, int tb);
// canon_in: the input may hold -0.0 (user data) and must go through "u0 + 0.0" literally;
// outputs of these kernels never hold -0.0, so later passes skip it (zero signs).
// Obstacle masks on the same path: pat = per-cell 5-bit neighbourhood pattern produced once from
// the mask by yh_rd_solid_patterns (nx*ny bytes, local rows); NULL = no mask.
int yh_rd_fast_solid_supported(const YhK &k, int tb);
int yh_rd_solid_patterns(const YhK &k, const uint8_t *solid, uint8_t *pat, cudaStream_t st);
int yh_launch_rd_fast(const YhK &k, int tb, const double *u_in, const double *v_in,
                      double *u_out, double *v_out, const uint8_t *pat, int canon_in,
                      cudaStream_t st);
int yh_launch_rd_fast_paced(const YhK &k, int tb, const double *u_in, const double *v_in,
                            double *u_out, double *v_out, int nsims, long long sim_stride,
                            const int *period_d, int duration_it, int count0, int canon_in,
                            cudaStream_t st, const YhApd *apd = nullptr, const uint8_t *pat = nullptr);

// Fused Runge-Kutta (RK2/RK4, optional 4th-order Laplacian) step, one launch per time step.
int yh_rd_rk_supported(const YhK &k);
int yh_launch_rd_rk(const YhK &k, const double *u_in, const double *v_in, double *u_out,
                    double *v_out, double *vtu, double *vtv, const uint8_t *solid, cudaStream_t st);

// Shared-memory tile kernels for small sheets (rd_tile.cu).
int yh_rd_prefer_tile(long long cells);
int yh_rd_tile_rk_supported(const YhK &k);
int yh_launch_rd_tile_rk(const YhK &k, const double *u_in, const double *v_in, double *u_out,
                         double *v_out, double *vtu, double *vtv, cudaStream_t st);
int yh_rd_tile_march_supported(const YhK &k);   // rd_tile_march.cu (YH_TILE_RK = march | cell overrides)
int yh_launch_rd_tile_march(const YhK &k, const double *u_in, const double *v_in, double *u_out,
                            double *v_out, double *vtu, double *vtv, cudaStream_t st);
int yh_rd_tile_march_solid_supported(const YhK &k);   // ... with obstacle masks (RK2 / RK4, exact flavour)
int yh_launch_rd_tile_march_solid(const YhK &k, const double *u_in, const double *v_in, double *u_out, double *v_out,
                                  double *vtu, double *vtv, const uint8_t *solid, cudaStream_t st);
// trace / slot (optional): the electrode (k.px, k.py) of every sheet is recorded at each of the tb
// levels into trace[2*((*slot + s)*nsims + z)]; yh_slot_bump advances the device-resident slot.
int yh_launch_rd_tile_euler(const YhK &k, int tb, const double *u_in, const double *v_in, double *u_out,
                            double *v_out, int nsims, long long sim_stride, const int *period_d,
                            int duration_it, int count0, cudaStream_t st, double *trace = nullptr,
                            const unsigned long long *slot = nullptr);
int yh_slot_bump(unsigned long long *slot, int n, cudaStream_t st);
