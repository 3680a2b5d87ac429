// abi.cu -- extern "C" entry points of libyolohtli_b200.so that are not kernels themselves:
// errors, workspace, parameter defaults, the RD step dispatch, the host-side 3x3 solve.
// There is NO CPU fallback anywhere in this library: without a CUDA device every compute
// entry point returns YH_ERR_NO_DEVICE.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include "yh_common.cuh"

static thread_local char g_err[512] = "";

void yh_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int yh_check_device(void) {
  static int cached = -1;
  if (cached < 0) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    cached = n;
  }
  if (cached == 0) {
    yh_set_error("no CUDA device visible: libyolohtli_b200 has no CPU fallback");
    return YH_ERR_NO_DEVICE;
  }
  return YH_OK;
}

// ---- workspace: a few growable slots per device ------------------------------------------
#define YH_WS_SLOTS 8
#define YH_MAX_DEV 32
static void *g_ws[YH_MAX_DEV][YH_WS_SLOTS];
static size_t g_ws_sz[YH_MAX_DEV][YH_WS_SLOTS];

int yh_workspace(size_t bytes, void **ptr, int slot) {
  int dev = 0;
  YH_CUDA(cudaGetDevice(&dev));
  if (dev >= YH_MAX_DEV || slot >= YH_WS_SLOTS) return YH_ERR_INVALID_ARG;
  if (g_ws_sz[dev][slot] < bytes) {
    if (g_ws[dev][slot]) {
      YH_CUDA(cudaDeviceSynchronize());
      YH_CUDA(cudaFree(g_ws[dev][slot]));
      g_ws[dev][slot] = nullptr; g_ws_sz[dev][slot] = 0;
    }
    YH_CUDA(cudaMalloc(&g_ws[dev][slot], bytes));
    YH_CUDA(cudaMemset(g_ws[dev][slot], 0, bytes));
    g_ws_sz[dev][slot] = bytes;
  }
  *ptr = g_ws[dev][slot];
  return YH_OK;
}

extern "C" {

int yh_abi_version(void) { return YH_ABI_VERSION; }
const char *yh_last_error(void) { return g_err; }

int yh_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int yh_release_workspace(void) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  int dev = 0;
  YH_CUDA(cudaGetDevice(&dev));
  YH_CUDA(cudaDeviceSynchronize());
  for (int s = 0; s < YH_WS_SLOTS; s++) {
    if (g_ws[dev][s]) YH_CUDA(cudaFree(g_ws[dev][s]));
    g_ws[dev][s] = nullptr; g_ws_sz[dev][s] = 0;
  }
  return YH_OK;
}

// saveFiles.cu:153-170, main.cu:149-155
int yh_params_derive(yh_params *p, double Dxx, double Dyy, double Dxy) {
  if (!p) return YH_ERR_INVALID_ARG;
  p->rxy = 2.0 * Dxy * p->dt / (4.0 * p->hx * p->hy);
  p->rbx = p->hx * Dxy / (Dxx * p->hy);
  p->rby = p->hy * Dxy / (Dyy * p->hx);
  p->rx = p->dt * Dxx / (p->hx * p->hx);
  p->ry = p->dt * Dyy / (p->hy * p->hy);
  p->invdx = 0.5 / p->hx;
  p->invdy = 0.5 / p->hy;
  p->qx4 = p->dt * Dyy / (p->hy * p->hy * 12.0);
  p->qy4 = p->dt * Dxx / (p->hx * p->hx * 12.0);
  p->fx4 = p->dt / 12.0;
  p->fy4 = p->dt / 12.0;
  return YH_OK;
}

// parameterSetup(), saveFiles.cu:105-231 + main.cu:148
int yh_params_default(yh_params *p, int nx, int ny, int reduce_sym, int scale_L) {
  if (!p || nx < 4 || ny < 4) { yh_set_error("yh_params_default: bad arguments"); return YH_ERR_INVALID_ARG; }
  memset(p, 0, sizeof(*p));
  p->nx = nx; p->ny = ny; p->ny_global = ny; p->jg0 = 0;
  p->solidSwitch = 0; p->neumannBC = 1; p->gateDiff = 1; p->anisotropy = 0;
  p->lap4 = 4; p->timeIntOrder = 4; p->tipGrad = 0; p->tipAlgorithm = 1;
  p->tipOffsetX = 160; p->tipOffsetY = 160; p->tipx0 = 0.0f; p->tipy0 = 0.0f;
  p->Lx = scale_L ? 12.0 * (nx - 1.0) / 511.0 : 12.0;
  p->Ly = scale_L ? 12.0 * (ny - 1.0) / 511.0 : 12.0;
  p->hx = p->Lx / (nx - 1.0);
  p->hy = p->Ly / (ny - 1.0);
  const double diff_par = 0.001, diff_per = 0.001, degrad = 0.0;
  const double th = degrad * 3.14159265359 / 180.0;
  const double Dxx = diff_par * cos(th) * cos(th) + diff_per * sin(th) * sin(th);
  const double Dyy = diff_par * sin(th) * sin(th) + diff_per * cos(th) * cos(th);
  const double Dxy = (diff_par - diff_per) * sin(th) * cos(th);
  p->rscale = 0.01;
  p->boundaryVal = 0.0; p->Uth = 0.7;
  p->tc = 1.0; p->alpha = 0.2; p->beta = 1.1; p->gamma = 0.0; p->delta = 1.0;
  p->eps = 0.005; p->mu = 1.0; p->theta = 0.0;
  p->dt = reduce_sym ? 0.5 * 0.02 : 0.02;
  return yh_params_derive(p, Dxx, Dyy, Dxy);
}

static int validate_rd(const yh_params *p, const void *a, const void *b, const void *c,
                       const void *d, const uint8_t *solid, int row0, int row1, int halo) {
  YH_REQUIRE(p && a && b && c && d, "null pointer");
  YH_REQUIRE(p->nx >= 4 && p->ny >= 1 && p->ny_global >= 4, "grid too small");
  YH_REQUIRE(p->jg0 >= 0 && p->jg0 + p->ny <= p->ny_global, "slab outside the global domain");
  YH_REQUIRE(!p->solidSwitch || solid, "solidSwitch set but solid == NULL");
  YH_REQUIRE(row0 >= 0 && row1 <= p->ny && row0 <= row1, "bad row range");
  // every row read must be stored locally: rows [row0-halo, row1+halo) clipped to the domain
  const int need_lo = row0 - halo > -p->jg0 ? row0 - halo : -p->jg0;
  const int dom_hi = p->ny_global - p->jg0;
  const int need_hi = row1 + halo < dom_hi ? row1 + halo : dom_hi;
  YH_REQUIRE(row0 == row1 || (need_lo >= 0 && need_hi <= p->ny), "not enough ghost rows for this row range");
  YH_REQUIRE(p->timeIntOrder == 1 || p->timeIntOrder == 2 || p->timeIntOrder == 4,
             "timeIntOrder must be 1, 2 or 4");
  return YH_OK;
}

int yh_rd_step(const yh_params *p, const double *u_in, const double *v_in, double *u_out,
               double *v_out, double *velTan_u, double *velTan_v, const uint8_t *solid,
               int stim_mouse, int point_x, int point_y, int row0, int row1, void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  rc = validate_rd(p, u_in, v_in, u_out, v_out, solid, row0, row1, p ? p->timeIntOrder : 1);
  if (rc != YH_OK) return rc;
  YH_REQUIRE(u_in != u_out && v_in != v_out, "in-place step is not supported (neighbours are read)");
  YH_REQUIRE((velTan_u == nullptr) == (velTan_v == nullptr), "velTan_u / velTan_v must both be set or NULL");
  YhK k = yh_make_k(p);
  k.stim = stim_mouse != 0; k.px = point_x; k.py = point_y; k.row0 = row0; k.row1 = row1;
  cudaStream_t st = (cudaStream_t)stream;
  const bool tile = yh_rd_prefer_tile((long long)k.nx * (row1 - row0)) != 0;   // small sheets
  // single steps that need no velTan take the Euler kernels (T = 1)
  if (!(velTan_u && k.gateDiff) && yh_rd_fast_supported(k, 1)) {
    if (tile) return yh_launch_rd_tile_euler(k, 1, u_in, v_in, u_out, v_out, 1, 0, nullptr, 0, 0, st);
    return yh_launch_rd_fast(k, 1, u_in, v_in, u_out, v_out, nullptr, 1, st);
  }
  if (tile && yh_rd_tile_rk_supported(k))
    return yh_launch_rd_tile_rk(k, u_in, v_in, u_out, v_out, velTan_u, velTan_v, st);
  if (yh_rd_rk_supported(k)) return yh_launch_rd_rk(k, u_in, v_in, u_out, v_out, velTan_u, velTan_v, solid, st);
  return yh_launch_rd_generic(k, u_in, v_in, u_out, v_out, velTan_u, velTan_v, solid, st);
}

int yh_rd_advance(const yh_params *p, int nsteps, int tb_steps, int flags, double *uA, double *vA,
                  double *uB, double *vB, const uint8_t *solid, int stim_mouse, int point_x,
                  int point_y, int row0, int row1, int *result_in_B, void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  YH_REQUIRE(nsteps >= 0 && result_in_B, "bad nsteps / result_in_B");
  YH_REQUIRE(tb_steps >= 0 && tb_steps <= 4 && tb_steps != 3, "tb_steps must be 0 (auto), 1, 2 or 4");
  rc = validate_rd(p, uA, vA, uB, vB, solid, row0, row1,
                   nsteps * (p ? p->timeIntOrder : 1));
  if (rc != YH_OK) return rc;
  YhK k = yh_make_k(p);
  k.stim = stim_mouse != 0; k.px = point_x; k.py = point_y;
  cudaStream_t st = (cudaStream_t)stream;
  double *cu = uA, *cv = vA, *nu = uB, *nv = vB;
  int inB = 0;
  const int K = p->timeIntOrder;
  const int dom_lo = -p->jg0, dom_hi = p->ny_global - p->jg0;
  int left = nsteps;
  int tb = tb_steps ? tb_steps : 4;
  int canon = (flags & YH_RD_INPUT_CANONICAL) ? 0 : 1;   // only the first pass can see user data
  // obstacle masks + Euler: the temporally blocked kernel reads per-cell mask patterns, made
  // once per call (mask contents belong to the caller and may change between calls)
  uint8_t *pat = nullptr;
  if (yh_rd_fast_solid_supported(k, 1) && nsteps > 1) {
    rc = yh_workspace((size_t)p->nx * p->ny, (void **)&pat, 4);
    if (rc != YH_OK) return rc;
    rc = yh_rd_solid_patterns(k, solid, pat, st);
    if (rc != YH_OK) return rc;
  }
  while (left > 0) {
    int T = 1;
    if (pat || yh_rd_fast_supported(k, 1)) { T = tb; while (T > left) T >>= 1; }
    // rows that must be valid after this pass so that the remaining steps stay exact
    const int ext = (left - T) * K;
    k.row0 = row0 - ext > dom_lo ? row0 - ext : dom_lo;
    k.row1 = row1 + ext < dom_hi ? row1 + ext : dom_hi;
    if (k.row0 < 0) k.row0 = 0;
    if (k.row1 > p->ny) k.row1 = p->ny;
    const bool tile = yh_rd_prefer_tile((long long)k.nx * (k.row1 - k.row0)) != 0;
    if (pat) {
      rc = yh_launch_rd_fast(k, T, cu, cv, nu, nv, pat, canon, st);
      canon = 0;
    } else if (yh_rd_fast_supported(k, T)) {
      if (tile) rc = yh_launch_rd_tile_euler(k, T, cu, cv, nu, nv, 1, 0, nullptr, 0, 0, st);
      else rc = yh_launch_rd_fast(k, T, cu, cv, nu, nv, nullptr, canon, st);
      canon = 0;
    } else if (tile && yh_rd_tile_rk_supported(k)) rc = yh_launch_rd_tile_rk(k, cu, cv, nu, nv, nullptr, nullptr, st);
    else if (yh_rd_rk_supported(k)) rc = yh_launch_rd_rk(k, cu, cv, nu, nv, nullptr, nullptr, solid, st);
    else rc = yh_launch_rd_generic(k, cu, cv, nu, nv, nullptr, nullptr, solid, st);
    if (rc != YH_OK) return rc;
    double *t = cu; cu = nu; nu = t;   // swapSoA (helper_functions.cu:140)
    t = cv; cv = nv; nv = t;
    inB ^= 1;
    left -= T;
  }
  *result_in_B = inB;
  return YH_OK;
}

// solve_matrix, symmetryReduction.cu:386-416 (host; same operation order; no leak)
int yh_solve_matrix(const double c_in[3], const double phi[3], const double Int[12],
                    double c_out[3]) {
  (void)c_in;
  if (!phi || !Int || !c_out) { yh_set_error("yh_solve_matrix: null pointer"); return YH_ERR_INVALID_ARG; }
  double a1, a2, a3, b1, b2, b3, C1, C2, C3, d1, d2, d3;
  double b2p, b3p, c2p, c3p, c3pp, d2p, d3p, d3pp, x1, x2, x3;
  const double pt = phi[2];
  a1 = Int[0] * cos(pt) + Int[1] * sin(pt); a2 = Int[1] * cos(pt) - Int[0] * sin(pt); a3 = Int[2];
  b1 = Int[3] * cos(pt) + Int[4] * sin(pt); b2 = Int[4] * cos(pt) - Int[3] * sin(pt); b3 = Int[5];
  C1 = Int[6] * cos(pt) + Int[7] * sin(pt); C2 = Int[7] * cos(pt) - Int[6] * sin(pt); C3 = Int[8];
  d1 = Int[9]; d2 = Int[10]; d3 = Int[11];
  b2p = a1 / b1 * b2 - a2;
  b3p = a1 / b1 * b3 - a3;
  d2p = a1 / b1 * d2 - d1;
  c2p = a1 / C1 * C2 - a2;
  c3p = a1 / C1 * C3 - a3;
  d3p = a1 / C1 * d3 - d1;
  c3pp = b2p / c2p * c3p - b3p;
  d3pp = b2p / c2p * d3p - d2p;
  x3 = d3pp / c3pp;
  x2 = (d2p - b3p * x3) / b2p;
  x1 = (d1 - a2 * x2 - a3 * x3) / a1;
  c_out[0] = x1; c_out[1] = x2; c_out[2] = x3;
  return YH_OK;
}

}  // extern "C"
