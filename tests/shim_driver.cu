// shim_driver.cu -- TEST: a miniature of the reference's display() loop (main.cu:869-1040)
// written against the reference's OWN launch API (hostPrototypes.h:22-57), linked with
// libyolohtli_shim.so instead of the reference's translation units.
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

#include "../include/yolohtli_abi.h"
#include "../include/yolohtli_compat.h"

extern "C" int yh_shimtest_run(const yh_params *p, double *u_h, double *v_h, int nsteps,
                               double *trace_h, double *velTan_u_h, vec5dyn *tips_h, int *ntips) {
  const size_t n = (size_t)p->nx * p->ny, bytes = n * sizeof(double);
  if (yh_shim_configure(p) != 0) return -1;
  size_t pitch = 0;
  dim3 grid2D((p->nx + 15) / 16, (p->ny + 15) / 16), block2D(16, 16), grid0D(1), block0D(1);
  stateVar gateIn_d, gateOut_d, J_d, velTan;
  cudaMalloc(&gateIn_d.u, bytes); cudaMalloc(&gateIn_d.v, bytes);
  cudaMalloc(&gateOut_d.u, bytes); cudaMalloc(&gateOut_d.v, bytes);
  cudaMalloc(&J_d.u, bytes); cudaMalloc(&J_d.v, bytes);
  cudaMalloc(&velTan.u, bytes); cudaMalloc(&velTan.v, bytes);
  bool *solid_d, *tip_plot; cudaMalloc(&solid_d, n); cudaMalloc(&tip_plot, n);
  cudaMemset(solid_d, 1, n); cudaMemset(tip_plot, 0, n);
  double *stimulus_d; cudaMalloc(&stimulus_d, bytes);
  int *tip_count_d; vec5dyn *tip_vector_d;
  cudaMalloc(&tip_count_d, sizeof(int)); cudaMalloc(&tip_vector_d, sizeof(vec5dyn) * 500000);
  double *point_d, point_h[2]; cudaMalloc(&point_d, 2 * sizeof(double));
  cudaMemcpy(gateIn_d.u, u_h, bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(gateIn_d.v, v_h, bytes, cudaMemcpyHostToDevice);
  const int2 point = make_int2(p->nx / 2, p->ny / 2);
  int count = 0;
  for (int i = 0; i < nsteps; i++) {
    reactionDiffusion_wrapper(pitch, grid2D, block2D, gateOut_d, gateIn_d, J_d, velTan, false, solid_d,
                              false, stimulus_d, false, point);
    swapSoA(&gateIn_d, &gateOut_d);
    count++;
    singleCell_wrapper(pitch, grid0D, block0D, gateOut_d, 2, point_h, point_d, point);
    trace_h[2 * i] = point_h[0]; trace_h[2 * i + 1] = point_h[1];
  }
  tip_wrapper(pitch, grid2D, block2D, gateIn_d, gateOut_d, velTan, p->dt * (double)count, p->tipAlgorithm,
              false, tip_plot, tip_count_d, tip_vector_d);
  cudaMemcpy(ntips, tip_count_d, sizeof(int), cudaMemcpyDeviceToHost);
  if (*ntips > 0) cudaMemcpy(tips_h, tip_vector_d, sizeof(vec5dyn) * (*ntips < 4096 ? *ntips : 4096), cudaMemcpyDeviceToHost);
  cudaMemcpy(u_h, gateIn_d.u, bytes, cudaMemcpyDeviceToHost);
  cudaMemcpy(v_h, gateIn_d.v, bytes, cudaMemcpyDeviceToHost);
  cudaMemcpy(velTan_u_h, velTan.u, bytes, cudaMemcpyDeviceToHost);
  double *f[] = {gateIn_d.u, gateIn_d.v, gateOut_d.u, gateOut_d.v, J_d.u, J_d.v, velTan.u, velTan.v, stimulus_d, point_d};
  for (double *q : f) cudaFree(q);
  cudaFree(solid_d); cudaFree(tip_plot); cudaFree(tip_count_d); cudaFree(tip_vector_d);
  return yh_shim_last_status();
}
