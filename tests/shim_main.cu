// shim_main.cu -- TEST: a host program built the way the reference's main.cu is: it DEFINES the global
// `paramVar param` (main.cu:37), fills it like parameterSetup() + main.cu:148-158 do, and drives a
// display()-style loop (main.cu:869-885, 1040) through the reference's own wrapper signatures
// (hostPrototypes.h:22-57).  It never calls yh_shim_configure: libyolohtli_shim.so picks the scalars up
// from `param` -- the zero-source-change link of SURVEY 8(b)(2).  The result is compared bit for bit with
// the same steps through the C ABI.
//
//   yh_shim_main [nx=256] [nsteps=60] [mode: default|euler]
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../include/yolohtli_abi.h"
#include "../include/yolohtli_compat.h"

paramVar param;   // the reference's global

int main(int argc, char **argv) {
  const int nx = argc > 1 ? atoi(argv[1]) : 256, nsteps = argc > 2 ? atoi(argv[2]) : 60;
  const bool euler = argc > 3 && strcmp(argv[3], "euler") == 0;
  if (yh_device_count() < 1) { printf("shim_main FAIL no CUDA device\n"); return 1; }
  yh_params p;
  if (yh_params_default(&p, nx, nx, 0, 0) != YH_OK) return 1;
  if (euler) { p.timeIntOrder = 1; p.lap4 = 0; }
  // what parameterSetup() leaves in the global (saveFiles.cu:105-231)
  memset(&param, 0, sizeof(param));
  param.nx = p.nx; param.ny = p.ny; param.solidSwitch = false; param.neumannBC = true; param.gateDiff = true;
  param.anisotropy = false; param.tipGrad = false; param.lap4 = p.lap4; param.timeIntOrder = p.timeIntOrder;
  param.tipAlgorithm = 1; param.Lx = p.Lx; param.Ly = p.Ly; param.hx = p.hx; param.hy = p.hy; param.dt = p.dt;
  param.rx = p.rx; param.ry = p.ry; param.rxy = p.rxy; param.rbx = p.rbx; param.rby = p.rby; param.rscale = p.rscale;
  param.invdx = p.invdx; param.invdy = p.invdy; param.qx4 = p.qx4; param.qy4 = p.qy4; param.fx4 = p.fx4; param.fy4 = p.fy4;
  param.boundaryVal = p.boundaryVal; param.tipOffsetX = p.tipOffsetX; param.tipOffsetY = p.tipOffsetY;
  param.Uth = p.Uth; param.tc = p.tc; param.alpha = p.alpha; param.beta = p.beta; param.gamma = p.gamma;
  param.delta = p.delta; param.eps = p.eps; param.mu = p.mu; param.theta = p.theta;
  param.contourThresh1 = 0.8; param.contourThresh2 = 0.85; param.contourThresh3 = 0.7;
  param.minVarColor = -0.1f; param.maxVarColor = 1.1f;

  const size_t n = (size_t)nx * nx, bytes = n * sizeof(double);
  std::vector<double> u0(n, 0.0), v0(n, 0.0), ua(n), va(n), ub(n), vb(n);
  for (int j = 0; j < nx; j++)          // initGates, main.cu:606-618
    for (int i = 0; i < nx; i++) {
      if (i < nx / 8) u0[(size_t)j * nx + i] = 1.0;
      if (j >= nx / 2) v0[(size_t)j * nx + i] = 1.0;
    }

  // (a) the reference's loop, linked against the shim
  size_t pitch = 0;
  dim3 grid2D((nx + 15) / 16, (nx + 15) / 16), block2D(16, 16), grid0D(1), block0D(1);
  stateVar gateIn_d, gateOut_d, J_d, velTan;
  cudaMalloc(&gateIn_d.u, bytes); cudaMalloc(&gateIn_d.v, bytes); cudaMalloc(&gateOut_d.u, bytes); cudaMalloc(&gateOut_d.v, bytes);
  cudaMalloc(&J_d.u, bytes); cudaMalloc(&J_d.v, bytes); cudaMalloc(&velTan.u, bytes); cudaMalloc(&velTan.v, bytes);
  bool *solid_d, *tip_plot;
  cudaMalloc(&solid_d, n); cudaMalloc(&tip_plot, n);
  cudaMemset(solid_d, 1, n);
  int *tip_count_d; vec5dyn *tip_vector_d;
  cudaMalloc(&tip_count_d, sizeof(int)); cudaMalloc(&tip_vector_d, sizeof(vec5dyn) * 500000);
  double *stimulus_d, *point_d, point_h[2];
  cudaMalloc(&stimulus_d, bytes); cudaMalloc(&point_d, 2 * sizeof(double));
  cudaMemcpy(gateIn_d.u, u0.data(), bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(gateIn_d.v, v0.data(), bytes, cudaMemcpyHostToDevice);
  const int2 point = make_int2(nx / 2, nx / 2);
  double trace = 0.0;
  for (int i = 0; i < nsteps; i++) {
    reactionDiffusion_wrapper(pitch, grid2D, block2D, gateOut_d, gateIn_d, J_d, velTan, false, solid_d, false,
                              stimulus_d, false, point);
    swapSoA(&gateIn_d, &gateOut_d);
    param.count++;
    param.physicalTime = param.dt * param.count;
    singleCell_wrapper(pitch, grid0D, block0D, gateOut_d, 2, point_h, point_d, point);
    trace += point_h[0];
  }
  tip_wrapper(pitch, grid2D, block2D, gateIn_d, gateOut_d, velTan, param.physicalTime, param.tipAlgorithm, false,
              tip_plot, tip_count_d, tip_vector_d);
  int ntips = -1;
  cudaMemcpy(&ntips, tip_count_d, sizeof(int), cudaMemcpyDeviceToHost);
  cudaMemcpy(ua.data(), gateIn_d.u, bytes, cudaMemcpyDeviceToHost);
  cudaMemcpy(va.data(), gateIn_d.v, bytes, cudaMemcpyDeviceToHost);
  if (yh_shim_last_status() != YH_OK) { printf("shim_main FAIL shim status %d: %s\n", yh_shim_last_status(), yh_last_error()); return 1; }

  // (b) the same steps through the C ABI
  yh_sim *sim = nullptr;
  if (yh_sim_create(&sim, &p, 1, 0) != YH_OK) { printf("shim_main FAIL %s\n", yh_last_error()); return 1; }
  if (yh_sim_run_host(sim, u0.data(), v0.data(), ub.data(), vb.data(), nsteps, 1) != YH_OK) { printf("shim_main FAIL %s\n", yh_last_error()); return 1; }
  yh_sim_destroy(sim);
  const bool same = memcmp(ua.data(), ub.data(), bytes) == 0 && memcmp(va.data(), vb.data(), bytes) == 0;
  double moved = 0.0;
  for (size_t c = 0; c < n; c++) moved += (ua[c] - u0[c]) * (ua[c] - u0[c]);
  const bool ok = same && moved > 1e-3 && ntips >= 0;
  printf("shim_main %s %dx%d %s nsteps=%d bitwise=%d tips=%d trace_sum=%.6f moved=%.3g (no yh_shim_configure call)\n",
         ok ? "PASS" : "FAIL", nx, nx, euler ? "euler" : "default", nsteps, (int)same, ntips, trace, moved);
  return ok ? 0 : 1;
}
