// slab_driver.cu -- TEST: a C++ host (no Python, no torch) that drives the monodomain step on N row
// slabs through include/yolohtli_slab.h, the way a maintainer of the reference's main.cu would reach
// several GPUs, and checks N slabs == one sheet BIT FOR BIT against yh_sim (single device path).
//
//   yh_slab_driver <nx> <ny> <nslabs> <nsteps> <mode: euler|rk4lap4|eulerholes|sr> [ndev] [pipe]
//
// sr: the symmetry-reduction branch of display() (main.cu:894-954) -- yh_slab_group_advance_sr on nslabs slabs
// against yh_sim_run_sr on one sheet: fields AND the (c, phi) record bit for bit, starting from a spiral that
// the program grows itself (cross-field initial condition, 3000 Euler steps).
//
// pipe: the slabs run through yh_slab_group_run_host (host buffers in, steps, host buffers out; the copies
// hidden behind the time steps by skewed chunks and edge wedges) in two calls; FAIL unless the pipelined
// schedule was really taken.
//
// Slab r lives on device r % ndev (ndev = 1: every slab on one GPU, the peers are then the same
// device and the flag protocol, streams and graphs are exercised exactly as across NVLink).
// Prints one line "slab_driver PASS ..." or "slab_driver FAIL ..."; exit status 0 / 1.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>

#include <chrono>
#include <vector>

#include "../include/yolohtli_slab.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    int rc__ = (call);                                                           \
    if (rc__ != YH_OK) {                                                         \
      printf("slab_driver FAIL %s -> %d: %s\n", #call, rc__, yh_last_error());   \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

// a few plane waves broken by a cross-field: something that moves everywhere, any size
static void initial_state(int nx, int ny, std::vector<double> &u, std::vector<double> &v) {
  for (int j = 0; j < ny; j++)
    for (int i = 0; i < nx; i++) {
      const size_t c = (size_t)j * nx + i;
      const int bx = (i * 8) / nx, by = (j * 8) / ny;
      u[c] = ((bx + by) % 3 == 0 && (i % 97) < 40) ? 1.0 : 0.0;
      v[c] = ((bx * 3 + by) % 4 == 0 && (j % 89) < 45) ? 0.9 : 0.05 * ((i + 2 * j) % 7);
    }
}

// symmetry-reduction mode: N slabs vs one sheet, fields and the (c, phi) history
static int run_sr(int nx, int ny, int nslabs, int nsteps, int ndev) {
  yh_params pe, p;
  // the spiral is grown on a sheet of at most 512 x 512 (the reference's default); larger sheets keep that grid
  // spacing (scale_L) and are tiled with copies of it, the integration disc sits on the copy nearest to the middle
  const int scale_L = nx > 512 || ny > 512;
  const int bx = nx > 512 ? 512 : nx, by = ny > 512 ? 512 : ny;
  if (nx % bx || ny % by) { printf("slab_driver FAIL sr: sheets larger than 512 must be multiples of 512\n"); return 1; }
  CHECK(yh_params_default(&pe, bx, by, 0, 0));
  pe.timeIntOrder = 1; pe.lap4 = 0;
  const size_t n = (size_t)nx * ny;
  std::vector<double> u0(n), v0(n), ua(n), va(n), ub(n), vb(n), reca((size_t)6 * nsteps), recb((size_t)6 * nsteps);
  yh_sim *sim = nullptr;
  {
    std::vector<double> us((size_t)bx * by), vs((size_t)bx * by);
    CHECK(yh_sim_create(&sim, &pe, 1, 0));
    CHECK(yh_sim_cross_field_ic(sim));
    CHECK(yh_sim_run(sim, 3000, 4, nullptr));
    CHECK(yh_sim_get_state(sim, us.data(), vs.data()));
    CHECK(yh_sim_destroy(sim));
    for (int j = 0; j < ny; j++)
      for (int i = 0; i < nx; i++) {
        u0[(size_t)j * nx + i] = us[(size_t)(j % by) * bx + i % bx];
        v0[(size_t)j * nx + i] = vs[(size_t)(j % by) * bx + i % bx];
      }
  }
  CHECK(yh_params_default(&p, nx, ny, 1, scale_L));      // reduce_sym: dt halved, default RK4 + lap4 (main.cu:148-158)
  p.tipx0 = (float)(bx / 2 + bx * ((nx / bx) / 2)); p.tipy0 = (float)(by / 2 + by * ((ny / by) / 2));
  if (p.tipOffsetX > bx / 3) { p.tipOffsetX = bx / 3; p.tipOffsetY = by / 3; }
  // (a) one sheet
  CHECK(yh_sim_create(&sim, &p, 1, 0));
  CHECK(yh_sim_set_state(sim, u0.data(), v0.data()));
  auto t0 = std::chrono::steady_clock::now();
  CHECK(yh_sim_run_sr(sim, nsteps, reca.data()));
  const double us_sheet = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / nsteps;
  CHECK(yh_sim_get_state(sim, ua.data(), va.data()));
  CHECK(yh_sim_destroy(sim));
  // (b) nslabs slabs, timeIntOrder + 3 ghost rows
  std::vector<int> devs(nslabs);
  for (int r = 0; r < nslabs; r++) devs[r] = r % ndev;
  yh_slab_group *g = nullptr;
  CHECK(yh_slab_group_create(&g, &p, nslabs, devs.data(), p.timeIntOrder + 3));
  CHECK(yh_slab_group_set_state(g, u0.data(), v0.data()));
  const int first = nsteps / 2;                     // in two calls: the run continues from resident state
  CHECK(yh_slab_group_advance_sr(g, first, recb.data()));
  t0 = std::chrono::steady_clock::now();
  CHECK(yh_slab_group_advance_sr(g, nsteps - first, recb.data() + (size_t)6 * first));
  const double us_slabs = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / (nsteps - first);
  CHECK(yh_slab_group_get_state(g, ub.data(), vb.data()));
  CHECK(yh_slab_group_destroy(g));
  const bool same = memcmp(ua.data(), ub.data(), n * sizeof(double)) == 0 && memcmp(va.data(), vb.data(), n * sizeof(double)) == 0;
  const bool same_rec = memcmp(reca.data(), recb.data(), reca.size() * sizeof(double)) == 0;
  double moved = 0.0, cmax = 0.0;
  for (size_t c = 0; c < n; c++) moved += (ua[c] - u0[c]) * (ua[c] - u0[c]);
  for (size_t q = 0; q < reca.size(); q++) cmax = reca[q] * reca[q] > cmax ? reca[q] * reca[q] : cmax;
  const bool ok = same && same_rec && moved > 1e-6 && cmax > 0.0 && cmax == cmax;
  printf("slab_driver %s %dx%d sr nslabs=%d ndev=%d nsteps=%d bitwise_fields=%d bitwise_c_phi=%d moved=%.3g max|c,phi|=%.3g "
         "us_per_step: sheet %.1f slabs %.1f\n",
         ok ? "PASS" : "FAIL", nx, ny, nslabs, ndev, nsteps, (int)same, (int)same_rec, moved, cmax > 0 ? sqrt(cmax) : 0.0,
         us_sheet, us_slabs);
  return ok ? 0 : 1;
}

int main(int argc, char **argv) {
  const int nx = argc > 1 ? atoi(argv[1]) : 512, ny = argc > 2 ? atoi(argv[2]) : 512;
  const int nslabs = argc > 3 ? atoi(argv[3]) : 2, nsteps = argc > 4 ? atoi(argv[4]) : 203;
  const char *mode = argc > 5 ? argv[5] : "euler";
  int ndev = argc > 6 ? atoi(argv[6]) : 1;
  const bool pipe = argc > 7 && strcmp(argv[7], "pipe") == 0;
  if (yh_device_count() < 1) { printf("slab_driver FAIL no CUDA device\n"); return 1; }
  if (ndev > yh_device_count()) ndev = yh_device_count();
  if (strcmp(mode, "sr") == 0) return run_sr(nx, ny, nslabs, nsteps, ndev);

  yh_params p;
  CHECK(yh_params_default(&p, nx, ny, 0, 1));
  const bool holes = strcmp(mode, "eulerholes") == 0;
  if (strncmp(mode, "euler", 5) == 0) { p.timeIntOrder = 1; p.lap4 = 0; }
  if (holes) p.solidSwitch = 1;
  const size_t n = (size_t)nx * ny;
  std::vector<double> u0(n), v0(n), ua(n), va(n), ub(n), vb(n);
  initial_state(nx, ny, u0, v0);
  std::vector<uint8_t> mask(n, 1);
  if (holes)
    for (int j = 0; j < ny; j++)
      for (int i = 0; i < nx; i++) {
        const int dx = i % 61 - 30, dy = j % 53 - 26;
        if (dx * dx + dy * dy < 90) { mask[(size_t)j * nx + i] = 0; u0[(size_t)j * nx + i] = 0.0; v0[(size_t)j * nx + i] = 0.0; }
      }

  // (a) one sheet, one device: the headless single-GPU driver
  yh_sim *sim = nullptr;
  CHECK(yh_sim_create(&sim, &p, 1, 0));
  if (holes) CHECK(yh_sim_set_solid(sim, mask.data()));
  CHECK(yh_sim_run_host(sim, u0.data(), v0.data(), ua.data(), va.data(), nsteps, 4));
  CHECK(yh_sim_destroy(sim));

  // (b) the same sheet on nslabs row slabs
  std::vector<int> devs(nslabs);
  for (int r = 0; r < nslabs; r++) devs[r] = r % ndev;
  yh_slab_group *g = nullptr;
  const int halo = 4;
  CHECK(yh_slab_group_create(&g, &p, nslabs, devs.data(), halo));
  if (holes) CHECK(yh_slab_group_set_solid(g, mask.data()));
  // in two calls, so that a run continues from device-resident state with valid ghosts, and with an
  // odd remainder so that the tail blocks (T = 2, 1) are exercised too
  const int first = nsteps / 3;
  int levels = -1;
  if (pipe) {
    std::vector<double> ut(n), vt(n);
    levels = yh_slab_pipeline_levels(yh_slab_group_member(g, 0), nsteps - first, 0);
    CHECK(yh_slab_group_run_host(g, u0.data(), v0.data(), ut.data(), vt.data(), first, 0));
    CHECK(yh_slab_group_run_host(g, ut.data(), vt.data(), ub.data(), vb.data(), nsteps - first, 0));
  } else {
    CHECK(yh_slab_group_set_state(g, u0.data(), v0.data()));
    CHECK(yh_slab_group_advance(g, first, 0));
    if (getenv("YH_SLAB_DRIVER_SYNC")) CHECK(yh_slab_group_sync(g));
    CHECK(yh_slab_group_advance(g, nsteps - first, 0));
    CHECK(yh_slab_group_get_state(g, ub.data(), vb.data()));
  }
  unsigned long long su = 0, sv = 0;
  for (int r = 0; r < nslabs; r++) {
    unsigned long long a = 0, b = 0;
    CHECK(yh_slab_checksum(yh_slab_group_member(g, r), &a, &b));
    su += a; sv += b;
  }
  CHECK(yh_slab_group_destroy(g));

  unsigned long long wu = 0, wv = 0;
  for (size_t c = 0; c < n; c++) {
    unsigned long long x, y;
    memcpy(&x, &ua[c], 8); memcpy(&y, &va[c], 8);
    wu += x; wv += y;
  }
  const bool same = memcmp(ua.data(), ub.data(), n * sizeof(double)) == 0 && memcmp(va.data(), vb.data(), n * sizeof(double)) == 0;
  double moved = 0.0;
  for (size_t c = 0; c < n; c++) moved += (ua[c] - u0[c]) * (ua[c] - u0[c]);
  const bool ok = same && su == wu && sv == wv && moved > 1e-3 && (!pipe || levels > 0);
  if (pipe) printf("slab_driver pipelined run_host: %d levels per chunk\n", levels);
  printf("slab_driver %s %dx%d %s nslabs=%d ndev=%d nsteps=%d bitwise=%d checksum=%016llx/%016llx (sheet %016llx/%016llx) moved=%.3g\n",
         ok ? "PASS" : "FAIL", nx, ny, mode, nslabs, ndev, nsteps, (int)same, su, sv, wu, wv, moved);
  return ok ? 0 : 1;
}
