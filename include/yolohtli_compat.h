/*
 * yolohtli_compat.h -- the handful of POD types the reference's launch API passes by value
 * (typeDefinition.cuh:4-20), re-declared so that libyolohtli_shim.so exports the SAME mangled
 * C++ symbols as the reference's own translation units (hostPrototypes.h:22-57).
 * Inside the reference tree, define YH_USE_REFERENCE_TYPES and include typeDefinition.cuh first.
 */
#ifndef YOLOHTLI_COMPAT_H
#define YOLOHTLI_COMPAT_H
#include <cuda_runtime.h>
#ifndef YH_USE_REFERENCE_TYPES
typedef double REAL;
typedef struct REAL3 { REAL x, y, t; } REAL3;
typedef struct stateVar { REAL *u, *v; } stateVar;
typedef struct advVar { REAL *x, *y; } advVar;
typedef struct sliceVar { REAL *ux, *uy, *ut, *vx, *vy, *vt; } sliceVar;
typedef struct vec5dyn { float x, y, vx, vy, t; } vec5dyn;
/* Layout of the reference's run-parameter block (typeDefinition.cuh:35-125).  The reference keeps ONE
 * global of this type, `paramVar param` (main.cu:37, `extern` in helper_functions.cu:17), filled by
 * parameterSetup() and main.cu:148-158 before the first wrapper call.  The shim reads the scalars of the
 * hot path from that global when the host program defines it, so main.cu links against
 * libyolohtli_shim.so WITHOUT SOURCE CHANGES (the cudaMemcpyToSymbol block of main.cu:309-402 then
 * fills __constant__ symbols nobody reads).  Field order and types must match the reference's. */
typedef struct paramVar {
  bool animate, saveEveryIt, plotTip, recordTip, plotContour, recordContour, stimulate, apdContour,
       plotTimeSeries, recordTimeSeries, reduceSym, reduceSymStart, clock, counterclock, firstIterTip,
       firstIterContour, firstFPS;
  bool solidSwitch, neumannBC, gateDiff, anisotropy, tipGrad;
  int lap4, contourMode, timeIntOrder, tipAlgorithm;
  bool load, save;
  int nx, ny, memSize;
  REAL Lx, Ly, hx, hy, dt, diff_par, diff_per, Dxx, Dyy, Dxy, rx, ry, rxy, rbx, rby, rscale, invdx, invdy, sample;
  int count;
  REAL physicalTime, physicalTimeLim;
  int startRecTime;
  REAL stimPeriod, stimDuration, stimMag;
  bool fibThreshold, fibTerminated;
  int leapShocks, eSize;
  int2 point;
  int nc;
  REAL rdomTrapz, rdomStim, rdomAPD, stcx, stcy;
  float2 pointStim;
  int savePackage;
  float tiempo;
  REAL degrad, boundaryVal, qx4, qy4, fx4, fy4;
  int itPerFrame, tipOffsetX, tipOffsetY;
  float minVarColor, maxVarColor;
  int wnx, wny;
  float uMax, uMin, vMax, vMin, tipx, tipy;
  REAL contourThresh1, contourThresh2, contourThresh3, Uth, tc, alpha, beta, gamma, delta, eps, mu, theta;
} paramVar;
#endif

/* The reference's wrappers (hostPrototypes.h:22-57), implemented by the shim. */
void reactionDiffusion_wrapper(size_t pitch, dim3 grid2D, dim3 block2D, stateVar gOut_d,
                               stateVar gIn_d, stateVar J, stateVar velTan, bool reduceSym,
                               bool *solid, bool stimLock, REAL *stim, bool stimLockMouse,
                               int2 point);
void tip_wrapper(size_t pitch, dim3 grid2D, dim3 block2D, stateVar gOut_d, stateVar gIn_d,
                 stateVar velTan, REAL physicalTime, int tipAlgorithm, bool recordTip,
                 bool *tip_plot, int *tip_count, vec5dyn *tip_vector);
void slice_wrapper(size_t pitch, dim3 grid2D, dim3 block2D, stateVar g, sliceVar slice,
                   sliceVar slice0, bool reduceSym, bool reduceSymStart, advVar adv, int scheme,
                   bool *intglArea, int *tip_count, vec5dyn *tip_vector, int count);
void Cxy_field_wrapper(size_t pitch, dim3 grid2D, dim3 block2D, advVar adv, REAL3 c, REAL3 phi,
                       bool *solid);
void advFDBFECC_wrapper(size_t pitch, dim3 grid2D, dim3 block2D, stateVar gOut, stateVar gIn,
                        advVar adv, stateVar uf, stateVar ub, stateVar ue, bool *solid);
REAL3 solve_matrix(REAL3 c, REAL3 phi, REAL *Int);
void trapz_wrapper(dim3 grid1D, dim3 block1D, sliceVar slice, sliceVar slice0, stateVar velTan,
                   REAL *integrals, REAL *coeffTrapz, int *tip_count, vec5dyn *tip_vector,
                   int count);
void singleCell_wrapper(size_t pitch, dim3 grid0D, dim3 block0D, stateVar gOut_d, int eSize,
                        REAL *pt_h, REAL *pt_d, int2 point);
void sAPD_wrapper(size_t pitch, dim3 grid1D, dim3 block1D, int count, REAL *uold, REAL *unew,
                  REAL *APD1, REAL *APD2, REAL *sAPD, REAL *dAPD, REAL *back, REAL *front,
                  bool *first, bool *stimArea, bool stimulate);
void countour_wrapper(size_t pitch, dim3 grid2D, dim3 block2D, REAL *field1, REAL *field2,
                      bool *contour_plot, bool *stimArea, int *contour_count,
                      float3 *contour_vector, float physicalTime, int mode);
void get_rgba_wrapper(size_t pitch, dim3 grid2D, dim3 block2D, int ncol, REAL *field,
                      unsigned int *plot_rba_data, unsigned int *cmap_rgba_data, bool *lines);
void swapSoA(stateVar *A, stateVar *B);

/* Explicit configuration for hosts that do not define the reference's global `paramVar param`
 * (replaces the ~45 cudaMemcpyToSymbol calls of main.cu:309-402).  When the host
 * program DOES define `param`, no call is needed: every wrapper reads it (weak reference). */
struct yh_params;
extern "C" int yh_shim_configure(const struct yh_params *p);
extern "C" int yh_shim_last_status(void);
/* conTh1_d..3 (main.cu:378-383) and minVarColor_d / maxVarColor_d (main.cu:366-369). */
extern "C" int yh_shim_set_contour_thresholds(double th1, double th2, double th3);
extern "C" int yh_shim_set_color_range(double min_var, double max_var);
#endif
