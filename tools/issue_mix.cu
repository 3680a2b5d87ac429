// issue_mix.cu -- does a non-FP64 instruction issue in the shadow of an FP64 instruction?
// B200's FP64 pipe takes a warp instruction every 2 cycles per scheduler (tools/fp64_peak.cu:
// 64 lanes/SM/clk).  This measures loops of 8 independent DADD/DMUL chains with N extra independent
// integer (IMAD/LOP), shared-memory (LDS) or select instructions per 8 FP64 ones: if the time
// stays flat up to N = 8, the other pipes ride in the FP64 shadow and a kernel is bound by
// max(2*FP64, all instructions); if it grows from N = 1, by 2*FP64 + others.
#include <cstdio>
#include <cuda_runtime.h>

template <int N, int KIND>
__global__ void __launch_bounds__(512) k(double *out, int iters, double a, double b, int ia) {
  __shared__ double sh[1024];
  double x[8];
  int y[16];
#pragma unroll
  for (int q = 0; q < 8; q++) x[q] = threadIdx.x * 1e-3 + q;
#pragma unroll
  for (int q = 0; q < 16; q++) y[q] = threadIdx.x + q;
  sh[threadIdx.x] = a; sh[threadIdx.x + 512] = b;
  __syncthreads();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
      x[q] = (q & 1) ? __dadd_rn(x[q], b) : __dmul_rn(x[q], a);
      if (q < N) {
        if (KIND == 0) y[q] = y[q] * ia + i;
        else if (KIND == 1) y[q] += __double2loint(sh[(threadIdx.x + y[q]) & 1023]);
        else y[q] = (y[q] & 4) ? y[q] + ia : i;
      }
      if (q + 8 < N) {
        if (KIND == 0) y[q + 8] = (y[q + 8] ^ i) + ia;
        else if (KIND == 1) y[q + 8] += __double2loint(sh[(threadIdx.x + y[q + 8]) & 1023]);
        else y[q + 8] = (y[q + 8] & 4) ? y[q + 8] + ia : i;
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 8; q++) s += x[q];
  int t = 0;
#pragma unroll
  for (int q = 0; q < 16; q++) t += y[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + t;
}

template <int N, int KIND>
void run(double *out, const char *kind) {
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  const int blocks = pr.multiProcessorCount * 2, threads = 512, iters = 100000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    k<N, KIND><<<blocks, threads>>>(out, iters, 0.999999, 1e-9, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double fp64 = (double)blocks * threads * iters * 8;
  printf("{\"kind\": \"%s\", \"extra_per_8_fp64\": %d, \"ms\": %.2f, \"fp64_Ginstr_per_s\": %.1f}\n", kind, N, best,
         fp64 / best / 1e6);
}

int main() {
  double *out;
  cudaMalloc(&out, sizeof(double) * 148 * 2 * 512);
  run<0, 0>(out, "int"); run<2, 0>(out, "int"); run<4, 0>(out, "int"); run<8, 0>(out, "int"); run<12, 0>(out, "int"); run<16, 0>(out, "int");
  run<2, 1>(out, "lds"); run<4, 1>(out, "lds"); run<8, 1>(out, "lds");
  run<2, 2>(out, "sel"); run<4, 2>(out, "sel"); run<8, 2>(out, "sel");
  return 0;
}
