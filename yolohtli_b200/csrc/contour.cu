// contour.cu -- contour extraction (countour_kernel modes 1-3, spaceAPD.cu:18-153, via
// countour_wrapper, :256-276) and the headless frame colouring (get_rgba_kernel,
// main.cu:1604-1641): SURVEY section 8 rows f1 and f4.
//
// Same pattern as the tip tracker: 16-24 B per cell of reads (HBM-bound), hits appended by the
// ORDERED compaction of yh_ordered.cuh, so the list comes out in ascending linear cell index
// (mode 3: the `-conTh2` crossing before the `+conTh3` crossing of a cell) instead of atomicAdd
// order, the count needs no cudaMemset, and one launch replaces {3 memsets, kernel}.
//
// Neighbour indexing as shipped: the reference reads I2D(i+1,j) and I2D(i,j+1) UNCLAMPED
// (:52-53, :78-79, :99-100).  For i = nx-1 that is the first cell of the next row -- reproduced,
// it is in bounds; for the last row (and the very last cell) it is past the end of the array:
// there the read returns the cell's own value (defect B13), which only matters where
// the reference's result is undefined.  plot_field (helper_functions.cu:45-51) keeps the shipped
// linear index floor(x) + nx*floor(y) (x = nx wraps into the next row) and is skipped only when
// that index is outside the array (the reference writes out of bounds).
#include "yh_common.cuh"
#include "yh_ordered.cuh"

namespace {

constexpr int CT_THREADS = 256;
constexpr int CT_CPT = 4;                        // consecutive cells per thread
constexpr int CT_CHUNK = CT_THREADS * CT_CPT;
constexpr double CT_EPS = 2.220446049250313e-16;   // MACHINE_EPS, globalVariables.cuh:26

struct ContourArgs {
  const double *f1, *f2;
  const uint8_t *stimArea;
  uint8_t *plot;
  int *count;
  yh_contour_pt *vec;
  int capacity, mode;
  float t;
  double th1, th2, th3;
  YhOrdered ord;
};

// Hits of one cell (0, 1 or 2); pts[] receives (x, y) of each.
__device__ __forceinline__ int contour_cell(const YhK &k, const ContourArgs &a, int i, int j,
                                            float2 *pts) {
  const int nx = k.nx, ny = k.ny;
  const long long c = (long long)i + (long long)nx * j;
  const long long ncell = (long long)nx * ny;
  const bool sc = a.stimArea ? a.stimArea[c] != 0 : true;
  const double f0 = a.f2[c];
  const double fx = (c + 1 < ncell) ? a.f2[c + 1] : f0;      // I2D(nx, i+1, j), unclamped in x
  const double fy = (c + nx < ncell) ? a.f2[c + nx] : f0;    // I2D(nx, i, j+1)
  int n = 0;
  if (a.mode == 1) {   // space-APD contour, :50-71
    const double v0 = f0, v1x = fx, v1y = fy;
    const double zpmx = i < (nx - 1) ? v0 * v1x : v0 * v0;
    const double zpmy = j < (ny - 1) ? v0 * v1y : v0 * v0;
    if ((zpmx < 0.0) || ((zpmy < 0.0) && sc)) {   // `||` binds looser than `&&`, as written
      const double ppx = fabs(v0 - v1x) > CT_EPS ? i + v0 / (v0 - v1x) : i;
      const double ppy = fabs(v0 - v1y) > CT_EPS ? j + v0 / (v0 - v1y) : j;
      pts[n++] = make_float2((float)ppx, (float)ppy);
    }
  } else if (a.mode == 2) {   // single contour, :74-95
    if ((a.f1[c] < a.th1) && sc) {
      const double v0 = f0 - a.th2, v1x = fx - a.th2, v1y = fy - a.th2;
      const double zpmx = i < (nx - 1) ? v0 * v1x : v0 * v0;
      const double zpmy = j < (ny - 1) ? v0 * v1y : v0 * v0;
      if ((zpmx < 0.0) || (zpmy < 0.0)) pts[n++] = make_float2((float)(double)i, (float)(double)j);
    }
  } else {   // double contour, :97-141
    const double V0 = f0 - a.th2, V1x = fx - a.th2, V1y = fy - a.th2;
    if ((a.f1[c] < a.th1) && (fabs(V0) < (a.th3 + 0.1)) && (fabs(V1x) < (a.th3 + 0.1)) &&
        (fabs(V1y) < (a.th3 + 0.1)) && sc) {
      double v0 = V0 - a.th2, v1x = V1x - a.th2, v1y = V1y - a.th2;
      double zpmx = i < (nx - 1) ? v0 * v1x : v0 * v0;
      double zpmy = j < (ny - 1) ? v0 * v1y : v0 * v0;
      if ((zpmx < 0.0) || (zpmy < 0.0)) pts[n++] = make_float2((float)(double)i, (float)(double)j);
      v0 = V0 + a.th3; v1x = V1x + a.th3; v1y = V1y + a.th3;
      zpmx = i < (nx - 1) ? v0 * v1x : v0 * v0;
      zpmy = j < (ny - 1) ? v0 * v1y : v0 * v0;
      if ((zpmx < 0.0) || (zpmy < 0.0)) pts[n++] = make_float2((float)(double)i, (float)(double)j);
    }
  }
  return n;
}

__global__ void __launch_bounds__(CT_THREADS)
contour_kernel(const __grid_constant__ YhK k, const __grid_constant__ ContourArgs a) {
  __shared__ int s_chunk, s_base;
  __shared__ int s_warp[CT_THREADS / 32];
  const int tid = threadIdx.x;
  if (tid == 0) s_chunk = (int)atomicAdd(&a.ord.state[0], 1ull);
  __syncthreads();
  const int chunk = s_chunk;
  const long long ncell = (long long)k.nx * k.ny;

  float2 pts[CT_CPT][2];
  int np[CT_CPT];
  int mine = 0;
#pragma unroll
  for (int r = 0; r < CT_CPT; r++) {
    const long long cell = (long long)chunk * CT_CHUNK + (long long)tid * CT_CPT + r;
    np[r] = 0;
    if (cell < ncell) np[r] = contour_cell(k, a, (int)(cell % k.nx), (int)(cell / k.nx), pts[r]);
    mine += np[r];
  }
  int total;
  const int excl = yh_block_excl_scan<CT_THREADS>(mine, s_warp, total);
  if (tid < 32) {   // warp 0 chains this chunk to its predecessors
    const unsigned prefix = yh_ordered_prefix(a.ord, chunk, (unsigned)total, a.count);
    if (tid == 0) s_base = (int)prefix;
  }
  __syncthreads();
  if (mine == 0) return;
  int pos = s_base + excl;
#pragma unroll
  for (int r = 0; r < CT_CPT; r++) {
    for (int q = 0; q < np[r]; q++, pos++) {
      yh_contour_pt d;
      d.x = pts[r][q].x; d.y = pts[r][q].y; d.t = a.t;
      if (pos < a.capacity) a.vec[pos] = d;
      if (a.plot) {   // plot_field, helper_functions.cu:45-51: I2D(nx, floor(x), floor(y)) as shipped
        const float fx = floorf(d.x), fy = floorf(d.y);
        if (fabsf(fx) < 1e9f && fabsf(fy) < 1e9f) {   // also rejects NaN / inf
          const long long idx = (long long)fx + (long long)k.nx * (long long)fy;
          if (idx >= 0 && idx < ncell) a.plot[idx] = 1;
        }
      }
    }
  }
}

// get_rgba_kernel, main.cu:1604-1631: icol = (int)((float)frac * (float)ncol), index as shipped
// except that it is clamped to [0, ncol-1] (the reference reads outside the colour map when the
// field leaves [minVarColor, maxVarColor)).
__global__ void rgba_kernel(long long n, const double *field, uint32_t *rgba, const uint32_t *cmap,
                            int ncol, double vmin, double vmax, const uint8_t *lines) {
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n;
       c += (long long)gridDim.x * blockDim.x) {
    const double frac = (field[c] - vmin) / (vmax - vmin);
    int icol = (int)((float)frac * (float)ncol);
    icol = icol < 0 ? 0 : (icol >= ncol ? ncol - 1 : icol);
    const uint32_t keep = lines ? (uint32_t)(!lines[c]) : 1u;
    rgba[c] = keep * cmap[icol];
  }
}

}  // namespace

extern "C" int yh_contour(const yh_params *p, const double *field1, const double *field2,
                          uint8_t *contour_plot, const uint8_t *stimArea, int *contour_count,
                          yh_contour_pt *contour_vector, int capacity, double physical_time,
                          int mode, double thresh1, double thresh2, double thresh3, void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  YH_REQUIRE(p && field2 && contour_count && contour_vector, "null pointer");
  YH_REQUIRE(mode >= 1 && mode <= 3, "contour mode must be 1, 2 or 3");
  YH_REQUIRE(mode == 1 || field1, "modes 2 and 3 need field1");
  YH_REQUIRE(capacity >= 0, "negative capacity");
  YH_REQUIRE(p->jg0 == 0 && p->ny_global == p->ny, "contour extraction works on a whole sheet");
  YhK k = yh_make_k(p);
  cudaStream_t st = (cudaStream_t)stream;
  const long long ncell = (long long)p->nx * p->ny;
  const int nchunks = (int)((ncell + CT_CHUNK - 1) / CT_CHUNK);
  unsigned long long *state = nullptr;
  rc = yh_workspace(((size_t)nchunks + 1) * sizeof(unsigned long long), (void **)&state, 3);
  if (rc != YH_OK) return rc;
  const unsigned epoch = yh_next_epoch();
  if (contour_plot) YH_CUDA(cudaMemsetAsync(contour_plot, 0, (size_t)ncell, st));   // :268
  ContourArgs a{field1, field2, stimArea, contour_plot, contour_count, contour_vector, capacity, mode,
                (float)physical_time, thresh1, thresh2, thresh3, {state, epoch, nchunks}};
  contour_kernel<<<nchunks, CT_THREADS, 0, st>>>(k, a);
  YH_LAUNCH_CHECK();
  return YH_OK;
}

extern "C" int yh_rgba(const yh_params *p, const double *field, uint32_t *plot_rgba,
                       const uint32_t *cmap_rgba, int ncol, double min_var, double max_var,
                       const uint8_t *lines, void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  YH_REQUIRE(p && field && plot_rgba && cmap_rgba && ncol > 0, "bad arguments");
  YH_REQUIRE(max_var != min_var, "empty colour range");
  const long long n = (long long)p->nx * p->ny;
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  rgba_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(n, field, plot_rgba, cmap_rgba, ncol,
                                                       min_var, max_var, lines);
  YH_LAUNCH_CHECK();
  return YH_OK;
}
