// apd.cu -- per-cell action-potential-duration bookkeeping (sAPD_kernel, spaceAPD.cu:278-374)
// and the electrode probe (singleCell_kernel, singleCell.cu:14-30).
#include "yh_common.cuh"

namespace {

struct ApdArgs {
  const double *uold, *unew;
  double *APD1, *APD2, *sAPD, *dAPD, *back, *front;
  uint8_t *first;
  const uint8_t *stimArea;
  int count, stimulate;
  long long n;
};

// One thread per cell; state is touched only where a crossing happens, so the steady-state
// traffic is the two field reads plus the sAPD (and dAPD) writes the reference also does.
__global__ void __launch_bounds__(256)
sapd_kernel(const __grid_constant__ YhK k, const __grid_constant__ ApdArgs a) {
  const double apdTh = 0.15;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < a.n;
       c += (long long)gridDim.x * blockDim.x) {
    const double uo = a.uold[c], un = a.unew[c];
    const bool sc = a.stimulate ? a.stimArea[c] != 0 : true;
    double fr = a.front[c], bk = a.back[c];
    bool dirty = false;
    if ((un > apdTh) && (uo < apdTh) && sc) { fr = k.dt * (a.count - (un - apdTh) / (un - uo)); dirty = true; }
    if ((un < apdTh) && (uo > apdTh) && sc) { bk = k.dt * (a.count - (un - apdTh) / (un - uo)); dirty = true; }
    bool first = a.first[c] != 0;
    double apd1 = a.APD1[c], apd2 = a.APD2[c];
    if ((bk > 0.0) && (fr > 0.0) && (first == false) && sc) {
      apd1 = bk - fr; a.APD1[c] = apd1; fr = 0.0; bk = 0.0; first = true; dirty = true;
    }
    if ((bk > 0.0) && (fr > 0.0) && first && sc) {
      apd2 = bk - fr; a.APD2[c] = apd2; fr = 0.0; bk = 0.0; first = false; dirty = true;
    }
    if (dirty) { a.front[c] = fr; a.back[c] = bk; a.first[c] = first ? 1 : 0; }
    if (a.stimulate) {
      double s = (apd1 - apd2 > 0.0) && sc ? 1.0 : -1.0;
      s *= (double)sc;
      a.sAPD[c] = s;
      double d = apd2;
      d *= (double)sc;
      a.dAPD[c] = d;
    } else {
      a.sAPD[c] = (apd1 - apd2 > 0.0) ? 1.0 : -1.0;
    }
  }
}

__global__ void probe_kernel(const double *u, const double *v, double *pt, long long idx) {
  pt[0] = u[idx];
  pt[1] = v[idx];
}

}  // namespace

extern "C" {

int yh_sapd(const yh_params *p, int count, const double *uold, const double *unew, double *APD1,
            double *APD2, double *sAPD, double *dAPD, double *back, double *front, uint8_t *first,
            const uint8_t *stimArea, int stimulate, void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  YH_REQUIRE(p && uold && unew && APD1 && APD2 && sAPD && back && front && first, "null pointer");
  YH_REQUIRE(!stimulate || (stimArea && dAPD), "stimulate needs stimArea and dAPD");
  YhK k = yh_make_k(p);
  ApdArgs a{uold, unew, APD1, APD2, sAPD, dAPD, back, front, first, stimArea, count, stimulate,
            (long long)p->nx * p->ny};
  long long blocks = (a.n + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  sapd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(k, a);
  YH_LAUNCH_CHECK();
  return YH_OK;
}

int yh_probe(const yh_params *p, const double *u, const double *v, double *pt_d, int x, int y,
             double *pt_h, void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  YH_REQUIRE(p && u && v && pt_d, "null pointer");
  YH_REQUIRE(x >= 0 && x < p->nx && y >= 0 && y < p->ny, "probe outside the sheet");
  cudaStream_t st = (cudaStream_t)stream;
  probe_kernel<<<1, 1, 0, st>>>(u, v, pt_d, (long long)x + (long long)p->nx * y);
  YH_LAUNCH_CHECK();
  if (pt_h) {   // the reference's blocking copy (singleCell.cu:28)
    YH_CUDA(cudaMemcpyAsync(pt_h, pt_d, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    YH_CUDA(cudaStreamSynchronize(st));
  }
  return YH_OK;
}

}  // extern "C"
