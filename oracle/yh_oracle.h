/*
 * yh_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of Yolohtli's 2D monodomain hot path (the "plain-C host loop"
 * named by BASELINE.json.north_star).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the product
 * (yolohtli_b200/) never links, imports or falls back to it.
 *
 * Semantics: SYNCHRONOUS stages.  The reference's RK/lap4/BFECC kernels
 * write stage values in place and neighbour-read them in the same launch
 * (reactionDiffusion.cu:117-121,219-229; advFDBFECC.cu:131-144), so their output depends on
 * block scheduling; here every neighbour read of stage k sees stage-k values.  In the
 * race-free modes (Euler, lap4=0) this is exactly the reference's arithmetic, expression by
 * expression, and is pinned BITWISE against the reference's own kernels built for sm_100
 * with --fmad=false (oracle/_ref, tests/test_ref_parity.py, fixtures under tests/golden/).
 * Racy modes: parity vs the reference is statistical only ("parity unpinned" beyond the
 * tolerance stated in DESIGN.md).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
 */
#ifndef YH_ORACLE_H
#define YH_ORACLE_H

#include "../include/yolohtli_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

int yho_params_default(yh_params *p, int nx, int ny, int reduce_sym, int scale_L);
int yho_params_derive(yh_params *p, double Dxx, double Dyy, double Dxy);

/* initGates cross-field IC, main.cu:606-618 */
void yho_cross_field_ic(int nx, int ny, double *u, double *v);

/* reactionDiffusion_kernel, reactionDiffusion.cu:26-566 (whole domain, jg0 must be 0). */
int yho_rd_step(const yh_params *p, const double *u_in, const double *v_in,
                double *u_out, double *v_out, double *velTan_u, double *velTan_v,
                const uint8_t *solid, int stim_mouse, int point_x, int point_y);
/* nsteps x {step; swap}: result copied to u/v in place (uses internal scratch). */
int yho_rd_advance(const yh_params *p, int nsteps, double *u, double *v,
                   const uint8_t *solid, int stim_mouse, int point_x, int point_y);
void yho_set_threads(int n);
int  yho_get_threads(void);

/* spiralTip / Newton / abouzar kernels + tipRecordPlane, tipTracker.cu:20-517.
 * Output in canonical order: ascending i + j*nx, '+' root before '-' root. */
int yho_tip_track(const yh_params *p, const double *u_past, const double *u_present,
                  uint8_t *tip_plot, int *tip_count, yh_tip *tip_vector, int capacity,
                  double physical_time, int algorithm);

/* slice_kernel, symmetryReduction.cu:72-222 */
int yho_slice(const yh_params *p, const double *u, const double *v,
              double *const slice[6], double *const slice0[6],
              int reduce_sym, int reduce_sym_start,
              const double *adv_x, const double *adv_y, int scheme,
              int tip_count, const yh_tip *tip_vector, int count);
/* trapz_kernel x12, integralTrapz.cu:17-184, canonical summation order (see .c). */
int yho_trapz(const yh_params *p, const double *const slice[6], const double *const slice0[6],
              const double *velTan_u, const double *velTan_v, double *integrals,
              int tip_count, const yh_tip *tip_vector, int count);
int yho_sr_integrals(const yh_params *p, const double *u, const double *v,
                     const double *velTan_u, const double *velTan_v,
                     const double *adv_x, const double *adv_y, double *integrals,
                     int tip_count, const yh_tip *tip_vector, int count);
/* solve_matrix, symmetryReduction.cu:386-416 */
int yho_solve_matrix(const double c_in[3], const double phi[3], const double Int[12],
                     double c_out[3]);
/* Cxy_field_kernel, symmetryReduction.cu:20-62 (host libm cos/sin) */
int yho_cxy_field(const yh_params *p, double *adv_x, double *adv_y,
                  const double c[3], const double phi[3], const uint8_t *solid);
/* advFDBFECC_kernel, advFDBFECC.cu:18-353, three synchronous sweeps */
int yho_advect_bfecc(const yh_params *p, const double *u_in, const double *v_in,
                     double *u_out, double *v_out,
                     const double *adv_x, const double *adv_y, const uint8_t *solid);
/* sAPD_kernel, spaceAPD.cu:278-374 */
int yho_sapd(const yh_params *p, int count, const double *uold, const double *unew,
             double *APD1, double *APD2, double *sAPD, double *dAPD,
             double *back, double *front, uint8_t *first, const uint8_t *stimArea,
             int stimulate);

/* countour_kernel modes 1-3 + countour_wrapper, spaceAPD.cu:18-153, 256-276 */
int yho_contour(const yh_params *p, const double *field1, const double *field2,
                uint8_t *contour_plot, const uint8_t *stimArea, int *contour_count,
                yh_contour_pt *contour_vector, int capacity, double physical_time, int mode,
                double thresh1, double thresh2, double thresh3);
/* get_rgba_kernel, main.cu:1604-1631 */
int yho_rgba(const yh_params *p, const double *field, uint32_t *plot_rgba, const uint32_t *cmap,
             int ncol, double vmin, double vmax, const uint8_t *lines);

#ifdef __cplusplus
}
#endif
#endif
