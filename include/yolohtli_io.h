/*
 * yolohtli_io.h -- host-side data formats either side of the hot path (SURVEY.md section 8f,
 * Appendix C): the files the reference reads and writes, plus a lossless binary snapshot.
 * Part of libyolohtli_b200.so; plain C ABI, HOST pointers only, no CUDA calls -- these entry
 * points work without a GPU (the reference's are host code too).  Every function returns a
 * YH_* status (yolohtli_abi.h); the reference's writers print and exit(0) on failure.
 *
 * Replaces (reference file:line):
 *   yh_io_run_params_default   parameterSetup                saveFiles.cu:105-231
 *   yh_io_params_write_csv     printParameters               printFunctions.cu:266-403
 *   yh_io_params_read_csv      loadParamValues               saveFiles.cu:540-711
 *   yh_io_state_write_text     print2D2column                printFunctions.cu:58-79
 *   yh_io_state_write_window   print2DSubWindow              printFunctions.cu:81-104
 *   yh_io_state_read_text      loadData                      saveFiles.cu:508-538
 *   yh_io_mask_read/_write     domainObjects mask parse      main.cu:676-680 (common/Hole_generator.m:27-40)
 *   yh_io_domain_objects       domainObjects (intglArea, stimArea, stimulus)   main.cu:686-848
 *   yh_io_tips_append          printTip                      printFunctions.cu:149-197
 *   yh_io_contour_append       printContour                  printFunctions.cu:199-247
 *   yh_io_sym_write            printSym                      printFunctions.cu:249-264
 *   yh_io_series_write         printVoltageInTime            printFunctions.cu:106-126
 *   yh_io_contour_length_write printContourLengthInTime      printFunctions.cu:128-147
 *   yh_io_cmap_read            loadcmap                      main.cu:1434-1470
 *   yh_io_reconstruct_tip      DATA/processSymmetry.m:68-89  (original-frame tip path)
 *   yh_io_snapshot_*           new: lossless FP64 checkpoint (the text formats keep 6 decimals)
 *   yh_io_frame_write_ppm      new: headless replacement of the PBO/GL frame (main.cu:1604-1641)
 */
#ifndef YOLOHTLI_IO_H
#define YOLOHTLI_IO_H

#include "yolohtli_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One contour point: layout of the float3 the reference appends (spaceAPD.cu:66). */
#ifndef YH_CONTOUR_PT_DEFINED
#define YH_CONTOUR_PT_DEFINED
typedef struct yh_contour_pt { float x, y, t; } yh_contour_pt;
#endif

/* The members of paramVar (typeDefinition.cuh:35-125) that dataparamcsv.csv carries, in the
 * file's own order.  `k` holds the kernel scalars; the rest is driver state. */
typedef struct yh_run_params {
  yh_params k;
  char read_path[200];       /* "Initial condition path:"  ("NA" when !load)              */
  char results_path[200];    /* "Results file path:"       ("NA" when !save)              */
  int32_t saveEveryIt, plotTip, recordTip, plotContour, recordContour, stimulate;
  int32_t plotTimeSeries, recordTimeSeries, reduceSym;
  int32_t contourMode, clock, counterclock;
  double diff_par, diff_per, degrad, Dxx, Dyy, Dxy;
  double physicalTimeLim, startRecTime;
  int32_t eSize, point_x, point_y;
  double stimPeriod, stimMag, stimDuration, fibThreshold;
  int32_t fibTerminated, leapShocks, nc;
  double stcx, stcy, rdomStim, rdomAPD, rdomTrapz;
  int32_t itPerFrame;
  double sample;
  double minVarColor, maxVarColor;
  int32_t wnx, wny;
  double uMin, uMax, vMin, vMax;
  double tipx, tipy;
  double contourThresh1, contourThresh2, contourThresh3;
} yh_run_params;

int yh_io_run_params_default(yh_run_params *rp, int nx, int ny);

/* dataparamcsv.csv: `label,value` lines, POSITIONAL (the labels are not parsed).  The writer
 * prints what printParameters prints (%d / %f, so doubles keep 6 decimals; the time step is
 * written doubled when reduceSym, as shipped).  The reader parses every value with strtof, like
 * the reference, assigns them in the same order, then recomputes the derived scalars
 * (rx..fy4) from (dt, hx, hy, Dxx, Dyy, Dxy) -- the reference reads rx.. back at 6 decimals. */
int yh_io_params_write_csv(const char *path, const yh_run_params *rp);
int yh_io_params_read_csv(const char *path, yh_run_params *rp);

/* raw_data.dat / dataSpiral.dat: one "%f %f\n" (u v) line per cell, i fastest.  Lossy (float,
 * 6 decimals).  The reader accepts any whitespace between the two numbers ("%f\t%f"). */
int yh_io_state_write_text(const char *path, const double *u, const double *v, int nx, int ny);
int yh_io_state_read_text(const char *path, double *u, double *v, int nx, int ny);
/* Window [floor(tip)-off-1, floor(tip)+off+1) around a tip; rows/columns outside the sheet are
 * skipped (the reference indexes out of bounds there).  *n_written = cells written. */
int yh_io_state_write_window(const char *path, const double *u, const double *v, int nx, int ny,
                             double tipx, double tipy, int offx, int offy, long long *n_written);

/* Lossless snapshot: 64-byte header {magic "YHSNAP01", nx, ny, n_sims, count, physical_time}
 * then u and v as raw little-endian FP64 (n_sims*nx*ny each). */
int yh_io_snapshot_write(const char *path, const double *u, const double *v, int nx, int ny,
                         int n_sims, long long count, double physical_time);
int yh_io_snapshot_info(const char *path, int *nx, int *ny, int *n_sims, long long *count,
                        double *physical_time);
int yh_io_snapshot_read(const char *path, double *u, double *v, long long capacity_cells);

/* Obstacle masks (holes<N>.dat, cBoundary<N>.dat): ASCII, n floats, value > 0.5 => tissue (1),
 * in file order = i + j*nx.  Bit-exact rule of main.cu:676-680. */
int yh_io_mask_read(const char *path, uint8_t *solid, long long n);
int yh_io_mask_write(const char *path, const uint8_t *solid, long long n);

/* The host-built masks of domainObjects (main.cu:686-848); any output may be NULL.  intglArea: disc of
 * radius rdomTrapz about the centre; stimArea: 1 where APD is measured (solid domains: outside the
 * disc rdomAPD about (stcx, stcy); square domains: rows j >= 35); stimulus: stimMag inside the
 * disc rdomStim about (stcx, stcy).  nx*ny entries each, i fastest. */
int yh_io_domain_objects(const yh_run_params *rp, uint8_t *intglArea, uint8_t *stimArea, double *stimulus);

/* dataTip.dat ("%f %f %f %f %f\n" = x y vx vy t) + dataTipSize.dat (count per sample; nothing is
 * written for an empty sample, as shipped).  first != 0 truncates both files. */
int yh_io_tips_append(const char *path_points, const char *path_counts, const yh_tip *tips,
                      int n, int first);
int yh_io_contour_append(const char *path_points, const char *path_counts,
                         const yh_contour_pt *pts, int n, int first);
/* c_phi_list_sym.dat: "cx cy ct phix phiy phit" per step; c_phi holds 6 doubles per step. */
int yh_io_sym_write(const char *path, const double *c_phi, int nsteps);
/* Electrode series: "t\te0\te1\t" per sample on ONE line, t = i*(float)dt*itPerFrame. */
int yh_io_series_write(const char *path, const double *e0, const double *e1, int n, double dt,
                       int itPerFrame);
int yh_io_contour_length_write(const char *path, const double *len, int n, double dt,
                               int itPerFrame);

/* Tip path in the original frame from the symmetry-reduced one (DATA/processSymmetry.m:68-89):
 *   xt = (x-1)*dx, yt = (y-1)*dy;  X = -phix - yt*sin(-phit) + xt*cos(-phit),
 *                                  Y = -phiy + xt*sin(-phit) + yt*cos(-phit).            */
int yh_io_reconstruct_tip(const float *tip_x, const float *tip_y, const double *c_phi, int n,
                          double dx, double dy, double *X, double *Y);

/* Colour map file: first the count, then "r g b" floats in [0,1]; packed as the reference does
 * (0xFF<<24 | b<<16 | g<<8 | r, components truncated from c*255.0f).  path == NULL fills a
 * built-in ramp of `capacity` entries. */
int yh_io_cmap_read(const char *path, uint32_t *cmap_rgba, int capacity, int *ncol);
/* Binary PPM (P6) of an nx x ny RGBA frame (row 0 at the bottom, like the GL window). */
int yh_io_frame_write_ppm(const char *path, const uint32_t *rgba, int nx, int ny);

#ifdef __cplusplus
}
#endif
#endif /* YOLOHTLI_IO_H */
