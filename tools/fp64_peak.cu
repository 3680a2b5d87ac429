// fp64_peak.cu -- measures the FP64 (non-tensor) DFMA and DADD/DMUL issue rate of the GPU,
// the second roofline of the monodomain step (RK4+lap4 is FP64-bound).
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    if (MODE == 0) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    } else {
      x0 = __dadd_rn(__dmul_rn(x0, a), b); x1 = __dadd_rn(__dmul_rn(x1, a), b);
      x2 = __dadd_rn(__dmul_rn(x2, a), b); x3 = __dadd_rn(__dmul_rn(x3, a), b);
      x4 = __dadd_rn(__dmul_rn(x4, a), b); x5 = __dadd_rn(__dmul_rn(x5, a), b);
      x6 = __dadd_rn(__dmul_rn(x6, a), b); x7 = __dadd_rn(__dmul_rn(x7, a), b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main() {
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  const int blocks = pr.multiProcessorCount * 4, threads = 512, iters = 200000;
  double *out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; mode++) {
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
      else k<1><<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double instr = (double)blocks * threads * iters * 8 * (mode == 0 ? 1 : 2);
      printf("{\"fp64_%s\": {\"Ginstr_per_s\": %.1f, \"TFLOPs\": %.2f, \"ms\": %.2f, \"sms\": %d}}\n",
             mode == 0 ? "dfma" : "dmul_dadd", instr / ms / 1e6, instr * (mode == 0 ? 2 : 1) / ms / 1e9, ms,
             pr.multiProcessorCount);
    }
  }
  return 0;
}
