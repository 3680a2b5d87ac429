/*
 * yolohtli_abi.h -- C ABI of the B200-native Yolohtli hot path (libyolohtli_b200.so).
 *
 * Every entry point is `extern "C"`, takes plain pointers / ints / doubles and a
 * CUDA stream as `void *` (0 = legacy default stream, which is what the reference
 * uses everywhere), and RETURNS AN INT STATUS (YH_OK == 0).  The reference's own
 * wrappers are `void` and only check errors under -DCUDA_ERROR_CHECK
 * (common/CudaSafeCall.h:4-41); the C++ re-export of those exact signatures lives in
 * yolohtli_b200/csrc/shim_wrappers.cu and forwards here.
 *
 * All field arrays are DEVICE pointers to dense, unpitched FP64 arrays, x fastest:
 * idx = i + j*nx   (reference: typeDefinition.cuh:4-9, reactionDiffusion.cu:47).
 * Masks are 1 byte per cell (C++ `bool`), true = tissue (main.cu:676-680).
 * Ownership: the caller allocates and frees every buffer it passes (main.cu:249-302);
 * the library owns only its internal workspace (released by yh_release_workspace()).
 *
 * Replaces (reference file:line):
 *   yh_configure / yh_params   ~45 __constant__ scalars      main.cu:40-51, 309-402
 *   yh_rd_step                 reactionDiffusion_wrapper      reactionDiffusion.cu:568-577
 *   yh_rd_advance              N x {reactionDiffusion_wrapper; swapSoA}   main.cu:879-882
 *   yh_tip_track               tip_wrapper                    tipTracker.cu:569-611
 *   yh_slice                   slice_wrapper                  symmetryReduction.cu:312-322
 *   yh_trapz                   trapz_wrapper (12 launches)    integralTrapz.cu:85-184
 *   yh_sr_integrals            slice_wrapper + trapz_wrapper fused, no slice arrays
 *   yh_cxy_field               Cxy_field_wrapper              symmetryReduction.cu:64-70
 *   yh_advect_bfecc            advFDBFECC_wrapper             advFDBFECC.cu:355-362
 *   yh_advect_bfecc_cphi       Cxy_field_wrapper + advFDBFECC_wrapper fused
 *   yh_solve_matrix            solve_matrix (host)            symmetryReduction.cu:329-420
 *   yh_sapd                    sAPD_wrapper                   spaceAPD.cu:376-384
 *   yh_probe                   singleCell_wrapper             singleCell.cu:22-30
 *   yh_contour                 countour_wrapper               spaceAPD.cu:256-276
 *   yh_rgba                    get_rgba_wrapper               main.cu:1633-1641
 *   yh_sim_*                   the display() step loop, headless   main.cu:862-1043
 */
#ifndef YOLOHTLI_ABI_H
#define YOLOHTLI_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YH_ABI_VERSION 1

/* status codes */
#define YH_OK                 0
#define YH_ERR_INVALID_ARG   -1   /* null pointer, bad dimension, unsupported flag value */
#define YH_ERR_CUDA          -2   /* a CUDA runtime call failed; see yh_last_error() */
#define YH_ERR_UNSUPPORTED   -3   /* mode combination the reference leaves undefined */
#define YH_ERR_CAPACITY      -4   /* tip list overflow (reference: printFunctions.cu:172-175) */
#define YH_ERR_NO_DEVICE     -5   /* no CUDA device: there is NO CPU fallback */

/* capacity of the tip list the reference allocates (typeDefinition.cuh: TIPVECSIZE) */
#define YH_TIPVECSIZE 500000

/* yh_rd_advance flags */
#define YH_RD_INPUT_CANONICAL 1
#define YH_RD_SOLID_IS_PATTERNS 2   /* `solid` holds the output of yh_rd_mask_patterns */

/* One record of the tip list: layout of the reference's vec5dyn (typeDefinition.cuh:19-20). */
typedef struct yh_tip {
  float x, y, vx, vy, t;
} yh_tip;

/*
 * Scalar parameters of the path.  Field names follow paramVar (typeDefinition.cuh:35-125)
 * and the __constant__ symbols of main.cu:40-51.  Filled by yh_params_default() with the
 * values of parameterSetup() (saveFiles.cu:105-231) + the derived block of main.cu:148-158.
 *
 * Slab fields (new; the reference is single-GPU): a process may hold only rows
 * [jg0, jg0+ny) of a global nx x ny_global domain, ghost rows included.  Mirror / Dirichlet
 * boundary rules apply at the GLOBAL edges only.  Single-domain use: ny_global = ny, jg0 = 0.
 */
typedef struct yh_params {
  int32_t nx;            /* cells per row                                             */
  int32_t ny;            /* rows held in the arrays passed to this process            */
  int32_t ny_global;     /* rows of the whole domain                                  */
  int32_t jg0;           /* global row index of local row 0                           */

  /* mode switches (saveFiles.cu:124-132) */
  int32_t solidSwitch;   /* 0/1 obstacle mask                                         */
  int32_t neumannBC;     /* 1 no-flux mirror, 0 Dirichlet boundaryVal                 */
  int32_t gateDiff;      /* 1: v diffuses (x rscale) and velTan is produced           */
  int32_t anisotropy;    /* cross-derivative term                                     */
  int32_t lap4;          /* truthy: 9-point 4th-order Laplacian + J correction        */
  int32_t timeIntOrder;  /* 1 Euler, 2 RK2 (midpoint), 4 RK4                          */
  int32_t tipGrad;       /* record u-gradient at tips                                 */
  int32_t tipAlgorithm;  /* 1 bilinear closed form, 2 Newton, 3 sign-change flag      */

  int32_t tipOffsetX, tipOffsetY;   /* integration-disc radius^2 = X*Y (cells)        */
  float   tipx0, tipy0;             /* disc centre used while count == 0              */

  double dt, hx, hy, Lx, Ly;
  double rx, ry, rxy, rbx, rby, rscale;
  double qx4, qy4, fx4, fy4;
  double invdx, invdy;
  double tc, alpha, beta, gamma, delta, eps, mu, theta;
  double boundaryVal, Uth;
} yh_params;

/* ---- library / errors ---------------------------------------------------------- */
int         yh_abi_version(void);
const char *yh_last_error(void);             /* thread-local text of the last failure   */
int         yh_device_count(void);           /* 0 => every compute call returns YH_ERR_NO_DEVICE */
int         yh_release_workspace(void);      /* frees internal scratch of this thread's device */

/* Arithmetic flavour of the reaction-diffusion kernels (process-wide; default YH_ARITH_EXACT, or the
 * environment variable YH_ARITH = exact | fast read at first use).
 *   EXACT  the reference's expressions operation for operation, no FMA contraction: bit-identical to
 *          the plain-C oracle and to the reference's kernels built with --fmad=false.
 *   FAST   the same update with the stencil coefficients combined on the host and FMA chains (the
 *          reference's shipped build contracts FMAs too, Makefile:9).  About half the FP64 instructions
 *          in the default RK4 + 4th-order-Laplacian mode.  Differs from EXACT by rounding only; the
 *          bounds are pinned in tests/test_gpu_arith.py.  Kernels without a FAST variant run EXACT. */
#define YH_ARITH_EXACT 0
#define YH_ARITH_FAST  1
int yh_set_arithmetic(int flavour);
int yh_get_arithmetic(void);

/* How the small-sheet Runge-Kutta kernel (csrc/rd_tile_march.cu) would cut `rows` x nx outputs into tiles on a device
 * with n_sm multiprocessors: tiling[0] = strip width (outputs per tile row, even, <= 64 - 2*stages), tiling[1] = band
 * height (<= 48 - 2*stages).  Host arithmetic, no device needed (tests/test_abi.py checks its invariants). */
int yh_rd_tile_march_tiling(int nx, int rows, int stages, int n_sm, int tiling[2]);

/* Defaults of parameterSetup() (saveFiles.cu:105-231) for an nx x ny grid with the
 * reference's hx (Lx = 12*(nx-1)/511 keeps hx at its 512-grid value when scale_L != 0).
 * Also applies main.cu:148-158 (dt halved when reduce_sym, rx..fy4 recomputed). */
int yh_params_default(yh_params *p, int nx, int ny, int reduce_sym, int scale_L);
/* Recompute rx, ry, rxy, rbx, rby, qx4, qy4, fx4, fy4, invdx, invdy from dt, hx, hy and
 * the diffusion tensor (Dxx, Dyy, Dxy) exactly as saveFiles.cu:153-170 + main.cu:148-155. */
int yh_params_derive(yh_params *p, double Dxx, double Dyy, double Dxy);

/* ---- (1) reaction-diffusion step ------------------------------------------------
 * One explicit step (u,v)^n -> (u,v)^{n+1} of reactionDiffusion_kernel with SYNCHRONOUS
 * RK stages (the reference's in-place stage writes race, reactionDiffusion.cu:117).
 * u_in/v_in are never written.  velTan_u/v may be NULL (not produced); the reference
 * always writes them when gateDiff (reactionDiffusion.cu:529-552).  solid may be NULL when
 * !solidSwitch.  stim_mouse/point_x/point_y: the live disc stimulus (reactionDiffusion.cu:54-61).
 * Rows written: local rows [row0, row1); pass row0=0,row1=p->ny for a whole array.        */
int yh_rd_step(const yh_params *p,
               const double *u_in, const double *v_in, double *u_out, double *v_out,
               double *velTan_u, double *velTan_v, const uint8_t *solid,
               int stim_mouse, int point_x, int point_y,
               int row0, int row1, void *stream);

/* nsteps steps, ping-ponging between (uA,vA) and (uB,vB); state starts in A.  Uses the
 * temporally-blocked kernels when the mode allows (Euler, 5-point), otherwise one pass per
 * step.  *result_in_B = 1 when the final state is in B.  Rows [row0,row1) are valid in the
 * result provided rows [row0-nsteps, row1+nsteps) (clipped to the global domain) were valid
 * in A -- the multi-GPU ghost-row contract.  tb_steps: time steps per HBM pass (0 = auto).
 * flags: YH_RD_INPUT_CANONICAL promises that A holds no -0.0 (true for anything this library
 * wrote); the first pass may then skip the literal "u0 + 0.0*0.0" of reactionDiffusion.cu:117. */
int yh_rd_advance(const yh_params *p, int nsteps, int tb_steps, int flags,
                  double *uA, double *vA, double *uB, double *vB,
                  const uint8_t *solid, int stim_mouse, int point_x, int point_y,
                  int row0, int row1, int *result_in_B, void *stream);

/* Obstacle masks + Euler: the temporally blocked kernel reads a per-cell neighbourhood pattern
 * (sc | sw<<1 | se<<2 | sn<<3 | ss<<4, the 32 cases of reactionDiffusion.cu:162-169) instead of
 * the mask.  yh_rd_advance derives it on every call (the mask belongs to the caller and may
 * change); a caller whose mask is fixed computes it once with yh_rd_mask_patterns (nx*ny bytes,
 * device) and passes it as `solid` with YH_RD_SOLID_IS_PATTERNS.  Euler + Neumann only: other
 * modes return YH_ERR_UNSUPPORTED for that flag. */
int yh_rd_mask_patterns(const yh_params *p, const uint8_t *solid, uint8_t *patterns, void *stream);

/* ---- (4) spiral-tip tracking -----------------------------------------------------
 * Resets *tip_count (device int) and appends every tip found between u_past and u_present.
 * The list is then put in canonical order (ascending linear cell index, '+' root before
 * '-' root) so that "the last tip" is deterministic.  tip_plot may be NULL.              */
int yh_tip_track(const yh_params *p, const double *u_past, const double *u_present,
                 uint8_t *tip_plot, int *tip_count, yh_tip *tip_vector, int capacity,
                 double physical_time, int algorithm, void *stream);

/* ---- (3) phase-condition integrals ----------------------------------------------
 * yh_slice fills the 12 tangent-field arrays exactly like slice_kernel.  slice[]/slice0[]
 * order: ux, uy, ut, vx, vy, vt (sliceVar, typeDefinition.cuh:15-17).  slice0 is written
 * only when reduce_sym_start.  Disc centre: (tipx0,tipy0) when count==0, else the last
 * entry of tip_vector (symmetryReduction.cu:98-104); an empty list keeps (tipx0,tipy0). */
int yh_slice(const yh_params *p, const double *u, const double *v,
             double *const slice[6], double *const slice0[6],
             int reduce_sym, int reduce_sym_start,
             const double *adv_x, const double *adv_y, int scheme,
             const int *tip_count, const yh_tip *tip_vector, int count, void *stream);

/* The 12 inner products of trapz_wrapper in ONE pass and ONE device->host copy.
 * integrals_host[12] is HOST memory (main.cu:220).  Deterministic summation order.     */
int yh_trapz(const yh_params *p, const double *const slice[6], const double *const slice0[6],
             const double *velTan_u, const double *velTan_v, double *integrals_host,
             const int *tip_count, const yh_tip *tip_vector, int count, void *stream);

/* slice + trapz fused: derivatives are formed in registers, the 12 slice arrays are never
 * materialised.  Same numbers as yh_slice(scheme 2, reduce_sym_start=1) + yh_trapz.      */
int yh_sr_integrals(const yh_params *p, const double *u, const double *v,
                    const double *velTan_u, const double *velTan_v,
                    const double *adv_x, const double *adv_y, double *integrals_host,
                    const int *tip_count, const yh_tip *tip_vector, int count, void *stream);

/* 3x3 elimination of solve_matrix (host, no pivoting, same operation order).           */
int yh_solve_matrix(const double c_in[3], const double phi[3], const double Int[12],
                    double c_out[3]);

/* ---- (2) co-moving-frame advection ----------------------------------------------- */
int yh_cxy_field(const yh_params *p, double *adv_x, double *adv_y,
                 const double c[3], const double phi[3], const uint8_t *solid, void *stream);

/* BFECC with three SYNCHRONOUS sweeps held in shared memory (no uf/ub/ue arrays).      */
int yh_advect_bfecc(const yh_params *p, const double *u_in, const double *v_in,
                    double *u_out, double *v_out,
                    const double *adv_x, const double *adv_y, const uint8_t *solid,
                    void *stream);

/* Same, with the advection field generated on the fly from (c, phi): Cxy fused.
 * adv_x/adv_y, when non-NULL, are also written (the next yh_slice reads their sign).    */
int yh_advect_bfecc_cphi(const yh_params *p, const double *u_in, const double *v_in,
                         double *u_out, double *v_out, const double c[3], const double phi[3],
                         double *adv_x, double *adv_y, const uint8_t *solid, void *stream);

/* ---- APD bookkeeping and electrode probe ------------------------------------------ */
int yh_sapd(const yh_params *p, int count, const double *uold, const double *unew,
            double *APD1, double *APD2, double *sAPD, double *dAPD,
            double *back, double *front, uint8_t *first, const uint8_t *stimArea,
            int stimulate, void *stream);

/* pt_d[0..1] = (u,v) at (x,y); no host sync.  pt_h non-NULL adds the reference's blocking copy. */
int yh_probe(const yh_params *p, const double *u, const double *v, double *pt_d,
             int x, int y, double *pt_h, void *stream);

/* ---- contour extraction and frame colouring --------------------------
 * One contour point: layout of the float3 the reference appends (spaceAPD.cu:66). */
#ifndef YH_CONTOUR_PT_DEFINED
#define YH_CONTOUR_PT_DEFINED
typedef struct yh_contour_pt { float x, y, t; } yh_contour_pt;
#endif
/* countour_kernel modes 1-3.  mode 1: field2 = sAPD (field1 unused, may be NULL); modes 2, 3:
 * field1 = u, field2 = v with the thresholds conTh1..3 (saveFiles.cu:215-217: 0.8, 0.85, 0.7).
 * Resets *contour_count (device int) and contour_plot (may be NULL), then appends every point
 * in canonical order (ascending linear cell index).  stimArea NULL = every cell counts.
 * *contour_count is the number found; at most `capacity` are stored. */
int yh_contour(const yh_params *p, const double *field1, const double *field2,
               uint8_t *contour_plot, const uint8_t *stimArea, int *contour_count,
               yh_contour_pt *contour_vector, int capacity, double physical_time, int mode,
               double thresh1, double thresh2, double thresh3, void *stream);
/* plot_rgba[c] = (!lines[c]) * cmap[(int)((float)frac*(float)ncol)], frac = (field-min)/(max-min);
 * the colour index is clamped to the map.  lines may be NULL. */
int yh_rgba(const yh_params *p, const double *field, uint32_t *plot_rgba,
            const uint32_t *cmap_rgba, int ncol, double min_var, double max_var,
            const uint8_t *lines, void *stream);

/* ---- headless driver: the display() loop without GL (main.cu:862-1043) ------------
 * An opaque simulation owning device state for n_sims independent nx x ny sheets
 * (batched parameter sweeps: sims are stacked along y and never exchange data).         */
typedef struct yh_sim yh_sim;

int yh_sim_create(yh_sim **out, const yh_params *p, int n_sims, int device);
int yh_sim_destroy(yh_sim *s);
/* Host <-> device state.  u_h/v_h hold n_sims*nx*ny doubles.                            */
int yh_sim_set_state(yh_sim *s, const double *u_h, const double *v_h);
int yh_sim_get_state(yh_sim *s, double *u_h, double *v_h);
int yh_sim_set_solid(yh_sim *s, const uint8_t *solid_h);       /* nx*ny bytes, shared by sims */
int yh_sim_cross_field_ic(yh_sim *s);                          /* initGates, main.cu:606-618 */
/* Advance nsteps standard-PDE steps.  Electrode trace: when trace_h != NULL it receives
 * 2*nsteps*n_sims doubles (u,v at p->point per step, one-step lag as main.cu:1040).     */
int yh_sim_set_point(yh_sim *s, int x, int y);
int yh_sim_run(yh_sim *s, int nsteps, int tb_steps, double *trace_h);
/* Per-sim pacing for sweeps: stimulus disc on for [k*period_it, k*period_it+duration_it]. */
int yh_sim_set_pacing(yh_sim *s, const int *period_it, int duration_it);
/* Whole reference use-case in one call with HOST buffers: H2D, nsteps, D2H.             */
int yh_sim_run_host(yh_sim *s, const double *u_in_h, const double *v_in_h,
                    double *u_out_h, double *v_out_h, int nsteps, int tb_steps);
int yh_sim_tips(yh_sim *s, yh_tip *tips_h, int capacity, int *count_out);
/* Symmetry-reduction (co-moving frame) steps, main.cu:894-954: RD -> tips -> phase-condition
 * integrals -> host 3x3 solve -> (Cxy + BFECC advection) -> phi += c*dt.  One host sync per step
 * (the 12 integrals).  c_phi_h (optional) receives 6 doubles per step: c then phi as pushed to
 * clist/philist (main.cu:902-903).  Sheet 0 only; (tipx0,tipy0) of the params centre the disc
 * while count == 0.  yh_sim_sr_state reads / sets (c, phi). */
int yh_sim_run_sr(yh_sim *s, int nsteps, double *c_phi_h);
/* Same steps with the integrals, the 3x3 solve and the frame update resident on the device: no
 * host round trip inside a step.  cos/sin(phi.t) are the device library's, so (c, phi) and the
 * fields agree with yh_sim_run_sr to rounding (1e-12 over hundreds of steps), not bit for bit. */
int yh_sim_run_sr_device(yh_sim *s, int nsteps, double *c_phi_h);
/* contourMode == 1 loop (main.cu:879-885, 1035): every step RD, swap, then sAPD_wrapper with the
 * reference's argument order (uold := gateIn = NEW state, unew := gateOut = OLD state) and
 * stimulate = 1 masked by stim_area_h (nx*ny bytes, uploaded and kept; NULL: the mask of an earlier
 * call, or -- none yet -- every cell, stimulate = 0).  All sheets
 * of the batch are processed by the same launches.  APD state is zero-initialised (defect B12).
 * yh_sim_get_apd copies APD1, APD2 (n_sims*nx*ny doubles each) to the host. */
int yh_sim_run_apd(yh_sim *s, int nsteps, const uint8_t *stim_area_h);
int yh_sim_get_apd(yh_sim *s, double *apd1_h, double *apd2_h);
int yh_sim_sr_state(yh_sim *s, double c[3], double phi[3], int set);
int yh_sim_count(const yh_sim *s);             /* param.count */
void *yh_sim_device_u(yh_sim *s);              /* current device pointers (for tests)     */
void *yh_sim_device_v(yh_sim *s);

/* ---- multi-GPU: row-slab forms of the symmetry-reduction pieces (new; SURVEY 8e, "SR mode") ---
 * A process holds rows [jg0, jg0+ny) of the nx x ny_global sheet (ghost rows included) and OWNS a
 * sub-range of them.  yh_slice, yh_cxy_field and yh_rd_step accept such a yh_params directly (all
 * coordinates are global, mirror rules apply at the global edges); the three passes whose whole-sheet
 * form ends in a global result have a slab form:
 *   tips       yh_tip_track_rows: the cells of local rows [row0,row1), GLOBAL coordinates in the list;
 *              lists of successive slabs concatenate to the whole-sheet list (ascending cell index).
 *   integrals  yh_sr_integral_rows: row sums of the 12 inner products for the owned disc rows ->
 *              rows_d[yh_sr_disc_slots(p)*12] (device; slots of rows owned elsewhere = +0.0).  The
 *              element-wise SUM over processes (ncclAllReduce; one non-zero contributor per slot, so
 *              exact) is closed by yh_sr_integrals_close into the 12 integrals on the host, bit for
 *              bit those of yh_sr_integrals on the whole sheet.  (centre_x, centre_y) = the last tip
 *              of the gathered list, or (tipx0, tipy0) when count == 0 / no tip.
 *   advection  yh_advect_bfecc_cphi_rows: local rows [row0,row1) written; needs 3 valid rows beyond.
 * Ghost depth of one SR step: timeIntOrder (RD) + 3 (BFECC); one exchange per step. */
int yh_tip_track_rows(const yh_params *p, const double *u_past, const double *u_present,
                      uint8_t *tip_plot, int *tip_count, yh_tip *tip_vector, int capacity,
                      double physical_time, int algorithm, int row0, int row1, void *stream);
int yh_sr_disc_slots(const yh_params *p);
int yh_sr_integral_rows(const yh_params *p, const double *u, const double *v,
                        const double *velTan_u, const double *velTan_v,
                        const double *adv_x, const double *adv_y,
                        float centre_x, float centre_y, int row0, int row1, double *rows_d,
                        void *stream);
int yh_sr_integrals_close(const yh_params *p, const double *rows_d, double *integrals_host,
                          void *stream);
int yh_advect_bfecc_cphi_rows(const yh_params *p, const double *u_in, const double *v_in,
                              double *u_out, double *v_out, const double c[3], const double phi[3],
                              double *adv_x, double *adv_y, const uint8_t *solid,
                              int row0, int row1, void *stream);

/* ---- multi-GPU: flags for the NVLink peer-to-peer halo exchange (new) -----------
 * yh_flag_set releases `value` into a flag word that may live in a PEER's memory (CUDA-IPC
 * mapping), stream-ordered after the ghost-row copies; yh_flag_wait blocks the stream until the
 * local flag reaches `value` (*status_local = 1 on a ~4 s timeout instead of hanging). */
int yh_flag_set(int *flag_peer, int value, void *stream);
int yh_flag_wait(int *flag_local, int value, int *status_local, void *stream);
/* cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault): dst may be a peer (CUDA-IPC) mapping. */
int yh_memcpy_async(void *dst, const void *src, size_t bytes, void *stream);
/* cudaDeviceEnablePeerAccess(peer_device) from the current device (already-enabled is not an error). */
int yh_enable_peer_access(int peer_device);

#ifdef __cplusplus
}
#endif
#endif /* YOLOHTLI_ABI_H */
