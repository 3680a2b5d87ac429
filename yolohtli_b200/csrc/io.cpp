// io.cpp -- the data formats either side of the hot path (include/yolohtli_io.h; SURVEY.md
// section 8 rows f2-f4, Appendix C).  Host code only: the reference's readers and writers are
// host code too (saveFiles.cu, printFunctions.cu, main.cu:1434-1470).  Formats are reproduced
// character for character where the reference writes them ("%f" of a float cast, tab / space
// separators, label text of dataparamcsv.csv); a lossless binary snapshot is added beside them.
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/yolohtli_io.h"

void yh_set_error(const char *fmt, ...);   // abi.cu

namespace {

struct File {
  FILE *f;
  File(const char *path, const char *mode) : f(path ? fopen(path, mode) : nullptr) {}
  ~File() { if (f) fclose(f); }
  operator FILE *() const { return f; }
};

int fail(const char *what, const char *path) {
  yh_set_error("%s: %s", what, path ? path : "(null path)");
  return YH_ERR_INVALID_ARG;
}

// ---- dataparamcsv.csv: one table drives both directions -----------------------------------
enum Kind { K_BOOL, K_INT, K_INT_SP, K_F, K_D, K_DT, K_PT };   // how printParameters prints the value

struct Row {
  const char *label;
  Kind kind;
  size_t off;   // offset in yh_run_params
};

#define RP(m) offsetof(yh_run_params, m)
#define KP(m) (offsetof(yh_run_params, k) + offsetof(yh_params, m))

// Rows in file order.  The derived scalars (rx.., qx4..) are printed from yh_params and, on the way
// back, overwritten by yh_params_derive.
const Row kRows[] = {
    // printFunctions.cu:293-310
    {"Save every sampling period:", K_BOOL, RP(saveEveryIt)},
    {"plot tip on screen:", K_BOOL, RP(plotTip)},
    {"Save tip in to data file:", K_BOOL, RP(recordTip)},
    {"Plot contours on screen:", K_BOOL, RP(plotContour)},
    {"Save contours to data file:", K_BOOL, RP(recordContour)},
    {"Pacing stimulus:", K_BOOL, RP(stimulate)},
    {"Plot time series:", K_BOOL, RP(plotTimeSeries)},
    {"Record time series:", K_BOOL, RP(recordTimeSeries)},
    {"Reduce symmetry:", K_BOOL, RP(reduceSym)},
    {"Solid boundary:", K_BOOL, KP(solidSwitch)},
    {"Neumann BCs:", K_BOOL, KP(neumannBC)},
    {"Gate diffusion:", K_BOOL, KP(gateDiff)},
    {"Tip trajectory algorithm:", K_INT_SP, KP(tipAlgorithm)},   // ", %d" with a space, :307
    {"Anisotropic tisue:", K_BOOL, KP(anisotropy)},
    {"Tip gradient:", K_BOOL, KP(tipGrad)},
    {"Contour mode (space APD or refractory):", K_INT, RP(contourMode)},
    {"Laplacian order:", K_BOOL, KP(lap4)},
    {"Scheme order for time integration (Euler or RK4):", K_INT, KP(timeIntOrder)},
    // :313-325
    {"Conduction block clock:", K_BOOL, RP(clock)},
    {"Conduction block counterclock:", K_BOOL, RP(counterclock)},
    {"# grid points X =", K_INT, KP(nx)},
    {"# grid points Y =", K_INT, KP(ny)},
    {"Physical Lx length", K_F, KP(Lx)},
    {"Physical Ly length", K_F, KP(Ly)},
    {"Physical dx", K_F, KP(hx)},
    {"Physical dy", K_F, KP(hy)},
    {"Time step:", K_DT, KP(dt)},
    // :327-350
    {"Diffusion parallel component:", K_F, RP(diff_par)},
    {"Diffusion perpendicular component:", K_F, RP(diff_per)},
    {"Initial fiber angle:", K_F, RP(degrad)},
    {"Diffusion Dxx:", K_F, RP(Dxx)},
    {"Diffusion Dyy:", K_F, RP(Dyy)},
    {"Diffusion Dxy:", K_F, RP(Dxy)},
    {"rxy (2*Dxy*dt/(4*dx*dy)):", K_F, KP(rxy)},
    {"rbx (hx*Dxy/(Dxx*dy)):", K_F, KP(rbx)},
    {"rby (hy*Dxy/(Dyy*dx)):", K_F, KP(rby)},
    {"rx (Dxx*dt/(dx*dx)):", K_F, KP(rx)},
    {"ry (Dyy*dt/(dy*dy)):", K_F, KP(ry)},
    {"Gate r-scale:", K_F, KP(rscale)},
    {"invdx (1/(2*hx))", K_F, KP(invdx)},
    {"invdy (1/(2*hy))", K_F, KP(invdy)},
    {"qx4 (dt*hx*hx*Dyy/12):", K_F, KP(qx4)},
    {"qy4 (dt*hy*hx*Dxx/12):", K_F, KP(qy4)},
    {"fx4 (dt*hx*hx/12):", K_F, KP(fx4)},
    {"fy4 (dt*hy*hy/12):", K_F, KP(fy4)},
    // :352-366
    {"Physical time limit:", K_F, RP(physicalTimeLim)},
    {"Start recording time:", K_F, RP(startRecTime)},
    {"Number of electrodes", K_INT, RP(eSize)},
    {"Electrode position x:", K_PT, RP(point_x)},
    {"Electrode position y:", K_PT, RP(point_y)},
    {"Stimulus period (ms):", K_D, RP(stimPeriod)},
    {"Stimulus magnitude:", K_F, RP(stimMag)},
    {"Stimulus duration (ms):", K_F, RP(stimDuration)},
    {"Voltage threshold for fibrillation:", K_F, RP(fibThreshold)},
    {"Fibrillation terminated by pacing:", K_BOOL, RP(fibTerminated)},
    {"Number of LEAP shocks applied:", K_INT, RP(leapShocks)},
    // :368-380
    {"Number of points in circles:", K_INT, RP(nc)},
    {"Stimulus position x:", K_F, RP(stcx)},
    {"Stimulus position y:", K_F, RP(stcy)},
    {"Stimulus area radius:", K_F, RP(rdomStim)},
    {"APD area radius:", K_F, RP(rdomAPD)},
    {"Tip offset x", K_INT, KP(tipOffsetX)},
    {"Tip offset y", K_INT, KP(tipOffsetY)},
    {"Integral area radius:", K_F, RP(rdomTrapz)},
    {"Dirichlet BC value:", K_F, KP(boundaryVal)},
    {"Iterations per frame:", K_INT, RP(itPerFrame)},
    {"Sampling period (ms):", K_D, RP(sample)},
    // :382-397
    {"Min signal range", K_F, RP(minVarColor)},
    {"Max signal range", K_F, RP(maxVarColor)},
    {"Secondary window size x", K_INT, RP(wnx)},
    {"Secondary window size y", K_INT, RP(wny)},
    {"Secondary window min signal 1", K_F, RP(uMin)},
    {"Secondary window max signal 1", K_F, RP(uMax)},
    {"Secondary window min signal 2", K_F, RP(vMin)},
    {"Secondary window max signal 2", K_F, RP(vMax)},
    {"Last tip point X", K_F, RP(tipx)},
    {"Last tip point Y", K_F, RP(tipy)},
    {"Contour threshold 1:", K_F, RP(contourThresh1)},
    {"Contour threshold 2:", K_F, RP(contourThresh2)},
    {"Contour threshold 3:", K_F, RP(contourThresh3)},
    {"Filament voltage threshold:", K_F, KP(Uth)},
    {"time scale (tc):", K_F, KP(tc)},
    {"alpha:", K_F, KP(alpha)},
    {"beta:", K_F, KP(beta)},
    {"gamma:", K_F, KP(gamma)},
    {"delta:", K_F, KP(delta)},
    {"epsilon:", K_F, KP(eps)},
    {"mu:", K_F, KP(mu)},
    {"theta:", K_F, KP(theta)},
};
constexpr int kNRows = (int)(sizeof(kRows) / sizeof(kRows[0]));

inline int32_t &as_i(yh_run_params *rp, size_t off) { return *(int32_t *)((char *)rp + off); }
inline double &as_d(yh_run_params *rp, size_t off) { return *(double *)((char *)rp + off); }
inline int32_t as_i(const yh_run_params *rp, size_t off) { return *(const int32_t *)((const char *)rp + off); }
inline double as_d(const yh_run_params *rp, size_t off) { return *(const double *)((const char *)rp + off); }

}  // namespace

extern "C" {

// parameterSetup(), saveFiles.cu:105-231 (the kernel scalars via yh_params_default)
int yh_io_run_params_default(yh_run_params *rp, int nx, int ny) {
  if (!rp) return fail("yh_io_run_params_default", "null");
  memset(rp, 0, sizeof(*rp));
  int rc = yh_params_default(&rp->k, nx, ny, 0, 0);
  if (rc != YH_OK) return rc;
  strcpy(rp->read_path, "NA");
  strcpy(rp->results_path, "NA");
  rp->saveEveryIt = 0; rp->plotTip = 1; rp->recordTip = 0; rp->plotContour = 0;
  rp->recordContour = 0; rp->stimulate = 0; rp->plotTimeSeries = 1; rp->recordTimeSeries = 0;
  rp->reduceSym = 0;
  rp->contourMode = 3; rp->clock = 0; rp->counterclock = 1;
  rp->diff_par = 0.001; rp->diff_per = 0.001; rp->degrad = 0.0;
  const double th = rp->degrad * 3.14159265359 / 180.0;
  rp->Dxx = rp->diff_par * cos(th) * cos(th) + rp->diff_per * sin(th) * sin(th);
  rp->Dyy = rp->diff_par * sin(th) * sin(th) + rp->diff_per * cos(th) * cos(th);
  rp->Dxy = (rp->diff_par - rp->diff_per) * sin(th) * cos(th);
  rp->physicalTimeLim = 100000.0; rp->startRecTime = 0.0;
  rp->eSize = 2; rp->point_x = nx / 2; rp->point_y = ny / 2;
  rp->stimPeriod = 600.0; rp->stimMag = 2.0; rp->stimDuration = 10.0; rp->fibThreshold = 0.1;
  rp->fibTerminated = 0; rp->leapShocks = 0;
  rp->nc = 100;
  rp->stcx = 0.25 * rp->k.Lx; rp->stcy = 0.25 * rp->k.Ly;
  rp->rdomStim = 0.03 * rp->k.Lx; rp->rdomAPD = 0.15 * rp->k.Lx;
  rp->rdomTrapz = 0.5 * ((rp->k.tipOffsetX + rp->k.tipOffsetY) * rp->k.hx);
  rp->itPerFrame = 50; rp->sample = 2.0;
  rp->minVarColor = -0.1f; rp->maxVarColor = 1.1f;
  rp->wnx = nx; rp->wny = ny;
  rp->uMin = -0.1f; rp->uMax = 1.1f; rp->vMin = -0.1f; rp->vMax = 0.5f;
  rp->tipx = 0.0; rp->tipy = 0.0;
  rp->contourThresh1 = 0.8; rp->contourThresh2 = 0.85; rp->contourThresh3 = 0.7;
  return YH_OK;
}

int yh_io_params_write_csv(const char *path, const yh_run_params *rp) {
  if (!rp) return fail("yh_io_params_write_csv", "null params");
  File f(path, "w+");
  if (!f) return fail("cannot create the parameter file", path);
  fprintf(f, "Initial condition path:,%s\n", rp->read_path[0] ? rp->read_path : "NA");
  fprintf(f, "Results file path:,%s\n", rp->results_path[0] ? rp->results_path : "NA");
  for (int r = 0; r < kNRows; r++) {
    const Row &w = kRows[r];
    switch (w.kind) {
      case K_BOOL: fprintf(f, "%s,%d\n", w.label, as_i(rp, w.off) ? 1 : 0); break;
      case K_INT: fprintf(f, "%s,%d\n", w.label, as_i(rp, w.off)); break;
      case K_INT_SP: fprintf(f, "%s, %d\n", w.label, as_i(rp, w.off)); break;
      case K_PT: fprintf(f, "%s,%f\n", w.label, (float)as_i(rp, w.off)); break;
      case K_F: fprintf(f, "%s,%f\n", w.label, (float)as_d(rp, w.off)); break;
      case K_D: fprintf(f, "%s,%f\n", w.label, as_d(rp, w.off)); break;
      case K_DT:   // printFunctions.cu:325: the halved SR step is written doubled
        fprintf(f, "%s,%f\n", w.label, rp->reduceSym ? 2.0 * (float)as_d(rp, w.off) : as_d(rp, w.off));
        break;
    }
  }
  return ferror(f) ? fail("write error", path) : YH_OK;
}

int yh_io_params_read_csv(const char *path, yh_run_params *rp) {
  if (!rp) return fail("yh_io_params_read_csv", "null params");
  File f(path, "r");
  if (!f) return fail("cannot open the parameter file", path);
  std::vector<std::string> second;   // text after the first comma of each line
  char line[512];
  while (fgets(line, sizeof(line), f)) {
    char *c = strchr(line, ',');
    std::string v = c ? std::string(c + 1) : std::string();
    while (!v.empty() && (v.back() == '\n' || v.back() == '\r')) v.pop_back();
    second.push_back(v);
  }
  if ((int)second.size() < 2 + kNRows) {
    yh_set_error("%s: %d lines, expected %d (positional format)", path, (int)second.size(), 2 + kNRows);
    return YH_ERR_INVALID_ARG;
  }
  snprintf(rp->read_path, sizeof(rp->read_path), "%s", second[0].c_str());
  snprintf(rp->results_path, sizeof(rp->results_path), "%s", second[1].c_str());
  for (int r = 0; r < kNRows; r++) {
    const Row &w = kRows[r];
    const float val = strtof(second[2 + r].c_str(), nullptr);   // saveFiles.cu:577
    switch (w.kind) {
      case K_BOOL: case K_INT: case K_INT_SP: case K_PT: as_i(rp, w.off) = (int)val; break;
      default: as_d(rp, w.off) = val; break;
    }
  }
  // saveFiles.cu:640 reads dt as written; main.cu:148 halves it again when reduceSym
  if (rp->reduceSym) rp->k.dt = 0.5 * rp->k.dt;
  rp->k.ny_global = rp->k.ny; rp->k.jg0 = 0;
  // main.cu:148-155 recomputes rx, ry, rxy, qx4, qy4, fx4, fy4 after loadParamValues and KEEPS rbx, rby, invdx,
  // invdy as the 6-decimal values read from the file; a restart here is bit-compatible with that
  yh_params *k = &rp->k;
  k->rx = k->dt * rp->Dxx / (k->hx * k->hx);
  k->ry = k->dt * rp->Dyy / (k->hy * k->hy);
  k->rxy = 2.0 * rp->Dxy * k->dt / (4.0 * k->hx * k->hy);
  k->qx4 = k->dt * rp->Dyy / (k->hy * k->hy * 12.0);
  k->qy4 = k->dt * rp->Dxx / (k->hx * k->hx * 12.0);
  k->fx4 = k->dt / 12.0;
  k->fy4 = k->dt / 12.0;
  rp->startRecTime = (double)(int)rp->startRecTime;   // saveFiles.cu:656: (int)
  return YH_OK;
}

// print2D2column, printFunctions.cu:58-79
int yh_io_state_write_text(const char *path, const double *u, const double *v, int nx, int ny) {
  if (!u || !v || nx <= 0 || ny <= 0) return fail("yh_io_state_write_text: bad arguments", path);
  File f(path, "w+");
  if (!f) return fail("cannot create the data file", path);
  const size_t n = (size_t)nx * ny;
  for (size_t c = 0; c < n; c++) fprintf(f, "%f %f\n", (float)u[c], (float)v[c]);
  return ferror(f) ? fail("write error", path) : YH_OK;
}

// loadData, saveFiles.cu:508-538 (values pass through float, as fscanf("%f") does)
int yh_io_state_read_text(const char *path, double *u, double *v, int nx, int ny) {
  if (!u || !v || nx <= 0 || ny <= 0) return fail("yh_io_state_read_text: bad arguments", path);
  File f(path, "r");
  if (!f) return fail("cannot open the initial condition file", path);
  const size_t n = (size_t)nx * ny;
  for (size_t c = 0; c < n; c++) {
    float a, b;
    if (fscanf(f, "%f%f", &a, &b) != 2) {
      yh_set_error("%s: short file, %zu of %zu cells", path, c, n);
      return YH_ERR_INVALID_ARG;
    }
    u[c] = a; v[c] = b;
  }
  return YH_OK;
}

// print2DSubWindow, printFunctions.cu:81-104
int yh_io_state_write_window(const char *path, const double *u, const double *v, int nx, int ny,
                             double tipx, double tipy, int offx, int offy, long long *n_written) {
  if (!u || !v || nx <= 0 || ny <= 0) return fail("yh_io_state_write_window: bad arguments", path);
  File f(path, "w+");
  if (!f) return fail("cannot create the data file", path);
  const int xmin = (int)floor(tipx) - offx - 1, xmax = (int)floor(tipx) + offx + 1;
  const int ymin = (int)floor(tipy) - offy - 1, ymax = (int)floor(tipy) + offy + 1;
  long long w = 0;
  for (int j = ymin; j < ymax; j++) {
    if (j < 0 || j >= ny) continue;
    for (int i = xmin; i < xmax; i++) {
      if (i < 0 || i >= nx) continue;
      const size_t c = (size_t)i + (size_t)nx * j;
      fprintf(f, "%f %f\n", (float)u[c], (float)v[c]);
      w++;
    }
  }
  if (n_written) *n_written = w;
  return ferror(f) ? fail("write error", path) : YH_OK;
}

// ---- lossless snapshot ---------------------------------------------------------------------
struct SnapHeader {
  char magic[8];
  int32_t nx, ny, n_sims, reserved;
  int64_t count;
  double physical_time;
  char pad[24];
};
static_assert(sizeof(SnapHeader) == 64, "snapshot header is 64 bytes");

int yh_io_snapshot_write(const char *path, const double *u, const double *v, int nx, int ny,
                         int n_sims, long long count, double physical_time) {
  if (!u || !v || nx <= 0 || ny <= 0 || n_sims <= 0) return fail("yh_io_snapshot_write: bad arguments", path);
  File f(path, "wb");
  if (!f) return fail("cannot create the snapshot", path);
  SnapHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, "YHSNAP01", 8);
  h.nx = nx; h.ny = ny; h.n_sims = n_sims; h.count = count; h.physical_time = physical_time;
  const size_t n = (size_t)nx * ny * n_sims;
  if (fwrite(&h, sizeof(h), 1, f) != 1 || fwrite(u, sizeof(double), n, f) != n ||
      fwrite(v, sizeof(double), n, f) != n)
    return fail("write error", path);
  return YH_OK;
}

static int snap_header(FILE *f, const char *path, SnapHeader *h) {
  if (fread(h, sizeof(*h), 1, f) != 1 || memcmp(h->magic, "YHSNAP01", 8) != 0 || h->nx <= 0 ||
      h->ny <= 0 || h->n_sims <= 0)
    return fail("not a YHSNAP01 snapshot", path);
  return YH_OK;
}

int yh_io_snapshot_info(const char *path, int *nx, int *ny, int *n_sims, long long *count,
                        double *physical_time) {
  File f(path, "rb");
  if (!f) return fail("cannot open the snapshot", path);
  SnapHeader h;
  int rc = snap_header(f, path, &h);
  if (rc != YH_OK) return rc;
  if (nx) *nx = h.nx;
  if (ny) *ny = h.ny;
  if (n_sims) *n_sims = h.n_sims;
  if (count) *count = h.count;
  if (physical_time) *physical_time = h.physical_time;
  return YH_OK;
}

int yh_io_snapshot_read(const char *path, double *u, double *v, long long capacity_cells) {
  if (!u || !v) return fail("yh_io_snapshot_read: null buffers", path);
  File f(path, "rb");
  if (!f) return fail("cannot open the snapshot", path);
  SnapHeader h;
  int rc = snap_header(f, path, &h);
  if (rc != YH_OK) return rc;
  const size_t n = (size_t)h.nx * h.ny * h.n_sims;
  if ((long long)n > capacity_cells) {
    yh_set_error("%s: %zu cells, buffer holds %lld", path, n, capacity_cells);
    return YH_ERR_CAPACITY;
  }
  if (fread(u, sizeof(double), n, f) != n || fread(v, sizeof(double), n, f) != n)
    return fail("truncated snapshot", path);
  return YH_OK;
}

// ---- masks: main.cu:676-680 -----------------------------------------------------------------
int yh_io_mask_read(const char *path, uint8_t *solid, long long n) {
  if (!solid || n <= 0) return fail("yh_io_mask_read: bad arguments", path);
  File f(path, "r");
  if (!f) return fail("cannot open the mask file", path);
  for (long long c = 0; c < n; c++) {
    float mesh;
    if (fscanf(f, "%f", &mesh) != 1) {
      yh_set_error("%s: short mask file, %lld of %lld values", path, c, n);
      return YH_ERR_INVALID_ARG;
    }
    solid[c] = mesh > 0.5 ? 1 : 0;
  }
  return YH_OK;
}

int yh_io_mask_write(const char *path, const uint8_t *solid, long long n) {
  if (!solid || n <= 0) return fail("yh_io_mask_write: bad arguments", path);
  File f(path, "w");
  if (!f) return fail("cannot create the mask file", path);
  for (long long c = 0; c < n; c++) fprintf(f, "%e\n", solid[c] ? 1.0 : 0.0);
  return ferror(f) ? fail("write error", path) : YH_OK;
}

// ---- domainObjects, main.cu:686-848: the host-built masks the kernels take ------------------
// intglArea (the disc of radius rdomTrapz about the domain centre, :696-703), stimArea (:816-848:
// solid domains exclude the disc of radius rdomAPD about the stimulus point, square domains the
// rows j < 35) and the stimulus field (:770-801).  x0 / y0 are `float` in the reference: the
// double expression is narrowed before the comparisons.  coeffTrapz is not built: no kernel reads
// it (integralTrapz.cu:30-32) and the reference's loop indexes intglArea[-1].
int yh_io_domain_objects(const yh_run_params *rp, uint8_t *intglArea, uint8_t *stimArea, double *stimulus) {
  if (!rp || rp->k.nx <= 0 || rp->k.ny <= 0) return fail("yh_io_domain_objects: bad arguments", nullptr);
  const yh_params &k = rp->k;
  const double rdomTrapz = 0.5 * ((k.tipOffsetX + k.tipOffsetY) * k.hx);   // recomputed at :691
  for (int j = 0; j < k.ny; j++) {
    for (int i = 0; i < k.nx; i++) {
      const size_t idx = (size_t)i + (size_t)k.nx * j;
      const float x0 = (float)((float)i * k.hx - 0.5 * k.Lx);
      const float y0 = (float)((float)j * k.hy - 0.5 * k.Ly);
      if (intglArea) intglArea[idx] = ((x0 * x0 + y0 * y0) < rdomTrapz * rdomTrapz) ? 1 : 0;
      const bool in_stim = ((x0 - rp->stcx) * (x0 - rp->stcx) + (y0 - rp->stcy) * (y0 - rp->stcy)) < rp->rdomStim * rp->rdomStim;
      if (stimulus) stimulus[idx] = in_stim ? rp->stimMag : 0.0;
      if (stimArea) {
        if (k.solidSwitch)
          stimArea[idx] = (((x0 - rp->stcx) * (x0 - rp->stcx) + (y0 - rp->stcy) * (y0 - rp->stcy)) < rp->rdomAPD * rp->rdomAPD) ? 0 : 1;
        else
          stimArea[idx] = (j < 35) ? 0 : 1;
      }
    }
  }
  return YH_OK;
}

// ---- series writers ---------------------------------------------------------------------------
// printTip, printFunctions.cu:149-197
int yh_io_tips_append(const char *path_points, const char *path_counts, const yh_tip *tips, int n,
                      int first) {
  if (n < 0 || (n > 0 && !tips)) return fail("yh_io_tips_append: bad arguments", path_points);
  if (n > YH_TIPVECSIZE) {
    yh_set_error("number of tip points (%d) exceeds the tip_vector size", n);
    return YH_ERR_CAPACITY;
  }
  File f1(path_points, first ? "w+" : "a+"), f2(path_counts, first ? "w+" : "a+");
  if (!f1 || !f2) return fail("cannot open the tip files", path_points);
  if (n > 0) {
    for (int i = 0; i < n; i++)
      fprintf(f1, "%f %f %f %f %f\n", tips[i].x, tips[i].y, tips[i].vx, tips[i].vy, tips[i].t);
    fprintf(f2, "%d\n", n);
  }
  return (ferror(f1) || ferror(f2)) ? fail("write error", path_points) : YH_OK;
}

// printContour, printFunctions.cu:199-247
int yh_io_contour_append(const char *path_points, const char *path_counts, const yh_contour_pt *pts,
                         int n, int first) {
  if (n < 0 || (n > 0 && !pts)) return fail("yh_io_contour_append: bad arguments", path_points);
  File f1(path_points, first ? "w+" : "a+"), f2(path_counts, first ? "w+" : "a+");
  if (!f1 || !f2) return fail("cannot open the contour files", path_points);
  if (n > 0) {
    for (int i = 0; i < n; i++) fprintf(f1, "%f %f %f\n", pts[i].x, pts[i].y, pts[i].t);
    fprintf(f2, "%d\n", n);
  }
  return (ferror(f1) || ferror(f2)) ? fail("write error", path_points) : YH_OK;
}

// printSym, printFunctions.cu:249-264
int yh_io_sym_write(const char *path, const double *c_phi, int nsteps) {
  if (nsteps < 0 || (nsteps > 0 && !c_phi)) return fail("yh_io_sym_write: bad arguments", path);
  File f(path, "w+");
  if (!f) return fail("cannot create the symmetry file", path);
  for (int i = 0; i < nsteps; i++) {
    const double *r = c_phi + 6 * (size_t)i;
    fprintf(f, "%f %f %f %f %f %f\n", r[0], r[1], r[2], r[3], r[4], r[5]);
  }
  return ferror(f) ? fail("write error", path) : YH_OK;
}

// printVoltageInTime, printFunctions.cu:106-126 (one line, tab separated, as shipped)
int yh_io_series_write(const char *path, const double *e0, const double *e1, int n, double dt,
                       int itPerFrame) {
  if (n < 0 || (n > 0 && (!e0 || !e1))) return fail("yh_io_series_write: bad arguments", path);
  File f(path, "w+");
  if (!f) return fail("cannot create the series file", path);
  for (int i = 0; i < n; i++) {
    fprintf(f, "%f\t", i * (float)dt * itPerFrame);
    fprintf(f, "%f\t", (float)e0[i]);
    fprintf(f, "%f\t", (float)e1[i]);
  }
  return ferror(f) ? fail("write error", path) : YH_OK;
}

// printContourLengthInTime, printFunctions.cu:128-147
int yh_io_contour_length_write(const char *path, const double *len, int n, double dt, int itPerFrame) {
  if (n < 0 || (n > 0 && !len)) return fail("yh_io_contour_length_write: bad arguments", path);
  File f(path, "w+");
  if (!f) return fail("cannot create the series file", path);
  for (int i = 0; i < n; i++) {
    fprintf(f, "%f\t", i * (float)dt * itPerFrame);
    fprintf(f, "%f\t", (float)len[i]);
  }
  return ferror(f) ? fail("write error", path) : YH_OK;
}

// DATA/processSymmetry.m:68-89
int yh_io_reconstruct_tip(const float *tip_x, const float *tip_y, const double *c_phi, int n,
                          double dx, double dy, double *X, double *Y) {
  if (n < 0 || (n > 0 && (!tip_x || !tip_y || !c_phi || !X || !Y)))
    return fail("yh_io_reconstruct_tip: bad arguments", nullptr);
  for (int i = 0; i < n; i++) {
    const double xt = ((double)tip_x[i] - 1.0) * dx, yt = ((double)tip_y[i] - 1.0) * dy;
    const double phix = c_phi[6 * (size_t)i + 3], phiy = c_phi[6 * (size_t)i + 4];
    const double phit = c_phi[6 * (size_t)i + 5];
    X[i] = (-phix - yt * sin(-phit) + xt * cos(-phit));
    Y[i] = (-phiy + xt * sin(-phit) + yt * cos(-phit));
  }
  return YH_OK;
}

// loadcmap, main.cu:1434-1470
int yh_io_cmap_read(const char *path, uint32_t *cmap_rgba, int capacity, int *ncol) {
  if (!cmap_rgba || capacity <= 0 || !ncol) return fail("yh_io_cmap_read: bad arguments", path);
  auto pack = [](float r, float g, float b) {
    return ((uint32_t)((int)(255.0f) << 24)) | ((uint32_t)((int)(b * 255.0f) << 16)) |
           ((uint32_t)((int)(g * 255.0f) << 8)) | ((uint32_t)((int)(r * 255.0f) << 0));
  };
  if (!path) {   // built-in ramp: dark teal -> yellow -> white (not the reference's data file)
    for (int i = 0; i < capacity; i++) {
      const float t = capacity > 1 ? (float)i / (float)(capacity - 1) : 0.f;
      const float r = t < 0.5f ? 2.f * t : 1.f, g = 0.35f + 0.65f * t, b = t < 0.5f ? 0.45f * (1.f - 2.f * t) : 2.f * t - 1.f;
      cmap_rgba[i] = pack(r, g, b);
    }
    *ncol = capacity;
    return YH_OK;
  }
  File f(path, "r");
  if (!f) return fail("cannot open the colour map", path);
  int n = 0;
  if (fscanf(f, "%d", &n) != 1 || n <= 0) return fail("bad colour map header", path);
  if (n > capacity) {
    yh_set_error("%s: %d colours, buffer holds %d", path, n, capacity);
    return YH_ERR_CAPACITY;
  }
  for (int i = 0; i < n; i++) {
    float r, g, b;
    if (fscanf(f, "%f%f%f", &r, &g, &b) != 3) return fail("short colour map", path);
    cmap_rgba[i] = pack(r, g, b);
  }
  *ncol = n;
  return YH_OK;
}

int yh_io_frame_write_ppm(const char *path, const uint32_t *rgba, int nx, int ny) {
  if (!rgba || nx <= 0 || ny <= 0) return fail("yh_io_frame_write_ppm: bad arguments", path);
  File f(path, "wb");
  if (!f) return fail("cannot create the frame", path);
  fprintf(f, "P6\n%d %d\n255\n", nx, ny);
  std::vector<unsigned char> row((size_t)nx * 3);
  for (int j = ny - 1; j >= 0; j--) {   // GL window: row 0 at the bottom
    for (int i = 0; i < nx; i++) {
      const uint32_t c = rgba[(size_t)i + (size_t)nx * j];
      row[3 * i + 0] = (unsigned char)(c & 0xFF);
      row[3 * i + 1] = (unsigned char)((c >> 8) & 0xFF);
      row[3 * i + 2] = (unsigned char)((c >> 16) & 0xFF);
    }
    if (fwrite(row.data(), 1, row.size(), f) != row.size()) return fail("write error", path);
  }
  return YH_OK;
}

}  // extern "C"
