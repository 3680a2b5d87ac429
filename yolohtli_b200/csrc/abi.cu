// abi.cu -- extern "C" entry points of libyolohtli_b200.so that are not kernels themselves:
// errors, workspace, parameter defaults, the RD step dispatch, the host-side 3x3 solve.
// There is NO CPU fallback anywhere in this library: without a CUDA device every compute
// entry point returns YH_ERR_NO_DEVICE.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "yh_common.cuh"

static thread_local char g_err[512] = "";
thread_local int yh_preload_only = 0;   // yh_common.cuh, YH_LAUNCH

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
static std::mutex g_ws_mutex;
// bumped whenever a slot is (re)allocated or freed: captured step loops hold workspace addresses, their
// cache key carries the generation they were captured under (every thread's cache goes stale at once)
static std::atomic<unsigned long long> g_ws_generation{1};
unsigned long long yh_workspace_generation(void) { return g_ws_generation.load(); }
// launch epochs of the ordered-compaction kernels (tip.cu, contour.cu): process-wide, so that two host
// threads sharing the per-device look-back words never produce the same epoch
unsigned yh_next_epoch(void) {
  static std::atomic<unsigned> e{0};
  unsigned v;
  do { v = (e.fetch_add(1) + 1) & 0xFFFFFFu; } while (v == 0);   // zero-initialised words must never look current
  return v;
}

void yh_graphs_release(void);

int yh_workspace(size_t bytes, void **ptr, int slot) {
  int dev = 0;
  YH_CUDA(cudaGetDevice(&dev));
  if (dev >= YH_MAX_DEV || slot >= YH_WS_SLOTS) return YH_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  if (g_ws_sz[dev][slot] < bytes) {
    g_ws_generation++;
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

static int g_arith = -1;
int yh_arithmetic(void) {
  if (g_arith < 0) {
    const char *f = getenv("YH_ARITH");
    g_arith = (f && (f[0] == 'f' || f[0] == 'F' || f[0] == '1')) ? YH_ARITH_FAST : YH_ARITH_EXACT;
  }
  return g_arith;
}

extern "C" {

int yh_set_arithmetic(int flavour) {
  if (flavour != YH_ARITH_EXACT && flavour != YH_ARITH_FAST) { yh_set_error("yh_set_arithmetic: unknown flavour"); return YH_ERR_INVALID_ARG; }
  g_arith = flavour;
  yh_graphs_release();   // captured step loops hold kernels of the other flavour
  return YH_OK;
}
int yh_get_arithmetic(void) { return yh_arithmetic(); }

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
  yh_graphs_release();   // captured graphs hold workspace addresses (other threads' caches: see the generation)
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  g_ws_generation++;
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
  rc = validate_rd(p, u_in, v_in, u_out, v_out, solid, row0, row1, p ? p->timeIntOrder * yh_rd_radius(p) : 1);
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

}  // extern "C"

// nsteps x {step; swapSoA} as plain launches on `st`; (cu,cv) holds the state on entry and on exit.
struct AdvanceCtx {
  const yh_params *p;
  YhK k;
  int tb, row0, row1;
  const uint8_t *solid, *pat;
};

static int advance_plain(const AdvanceCtx &x, int nsteps, int &canon, double *&cu, double *&cv,
                         double *&nu, double *&nv, int &inB, cudaStream_t st) {
  const yh_params *p = x.p;
  YhK k = x.k;
  const int K = p->timeIntOrder;
  const int dom_lo = -p->jg0, dom_hi = p->ny_global - p->jg0;
  int left = nsteps;
  while (left > 0) {
    int T = 1;
    if (x.pat || yh_rd_fast_supported(k, 1)) { T = x.tb; while (T > left) T >>= 1; }
    // rows that must be valid after this pass so that the remaining steps stay exact
    const int ext = (left - T) * K * yh_rd_radius(p);
    k.row0 = x.row0 - ext > dom_lo ? x.row0 - ext : dom_lo;
    k.row1 = x.row1 + ext < dom_hi ? x.row1 + ext : dom_hi;
    if (k.row0 < 0) k.row0 = 0;
    if (k.row1 > p->ny) k.row1 = p->ny;
    const bool tile = yh_rd_prefer_tile((long long)k.nx * (k.row1 - k.row0)) != 0;
    int rc;
    if (x.pat) {
      rc = yh_launch_rd_fast(k, T, cu, cv, nu, nv, x.pat, canon, st);
      canon = 0;
    } else if (yh_rd_fast_supported(k, T)) {
      if (tile) rc = yh_launch_rd_tile_euler(k, T, cu, cv, nu, nv, 1, 0, nullptr, 0, 0, st);
      else rc = yh_launch_rd_fast(k, T, cu, cv, nu, nv, nullptr, canon, st);
      canon = 0;
    } else if (tile && yh_rd_tile_rk_supported(k)) rc = yh_launch_rd_tile_rk(k, cu, cv, nu, nv, nullptr, nullptr, st);
    else if (yh_rd_rk_supported(k)) rc = yh_launch_rd_rk(k, cu, cv, nu, nv, nullptr, nullptr, x.solid, st);
    else rc = yh_launch_rd_generic(k, cu, cv, nu, nv, nullptr, nullptr, x.solid, st);
    if (rc != YH_OK) return rc;
    double *t = cu; cu = nu; nu = t;   // swapSoA (helper_functions.cu:140)
    t = cv; cv = nv; nv = t;
    inB ^= 1;
    left -= T;
  }
  return YH_OK;
}

// ---- CUDA-graph replay of the step loop for small sheets ------------------------------------
// A 512^2 step is a few microseconds of kernel time; dependent launches in a stream complete on a
// ~2 us cadence (measured: wall time per launch is a multiple of 2.05 us, profiles/), which a
// captured graph does not pay.  A chunk of YH_GRAPH_CHUNK time steps (an even number of launches,
// so the ping-pong returns to the same buffers) is captured once per {parameters, buffers} and
// replayed; the head of the run (raw input, remainder) goes through plain launches.
// Capture needs a non-legacy stream: graphs run on an internal stream fenced with events against
// the caller's stream, so the call stays stream-ordered for the caller.
#define YH_GRAPH_CHUNK 64
#define YH_GRAPH_CACHE 8

struct GraphKey {   // compared with memcmp: no padding anywhere (asserted below), zeroed before it is filled
  yh_params p;
  int stim, px, py, tb, row0, row1;
  const void *cu, *cv, *nu, *nv, *solid, *pat;
  unsigned long long ws_generation;
  long long arith;
};
static_assert(sizeof(yh_params) == 16 * 4 + 27 * 8, "yh_params must not contain padding");
static_assert(sizeof(GraphKey) == sizeof(yh_params) + 6 * 4 + 6 * 8 + 16, "GraphKey must not contain padding");
struct GraphEntry {
  GraphKey key;
  cudaGraphExec_t exec;
  unsigned long long stamp;
  int dev;
};
static thread_local GraphEntry g_graphs[YH_GRAPH_CACHE];
static thread_local unsigned long long g_graph_clock = 0;
static cudaStream_t g_graph_stream[YH_MAX_DEV];
static cudaEvent_t g_graph_ev[YH_MAX_DEV][2];

static int graph_stream(int dev, cudaStream_t *gs) {
  if (!g_graph_stream[dev]) {
    YH_CUDA(cudaStreamCreateWithFlags(&g_graph_stream[dev], cudaStreamNonBlocking));
    YH_CUDA(cudaEventCreateWithFlags(&g_graph_ev[dev][0], cudaEventDisableTiming));
    YH_CUDA(cudaEventCreateWithFlags(&g_graph_ev[dev][1], cudaEventDisableTiming));
  }
  *gs = g_graph_stream[dev];
  return YH_OK;
}

int yh_graphs_enabled(long long cells) {
  const char *f = getenv("YH_GRAPHS");   // 0: never, 1: whenever the shape allows
  if (f && f[0] == '0') return 0;
  if (f && f[0] == '1') return 1;
  return cells <= (3ll << 20);           // the launch-cadence regime: small sheets only
}

static GraphKey graph_key(const AdvanceCtx &x, const double *cu, const double *cv, const double *nu,
                          const double *nv) {
  GraphKey key;
  memset(&key, 0, sizeof(key));
  key.p = *x.p; key.stim = x.k.stim; key.px = x.k.px; key.py = x.k.py; key.tb = x.tb;
  key.row0 = x.row0; key.row1 = x.row1;
  key.cu = cu; key.cv = cv; key.nu = nu; key.nv = nv; key.solid = x.solid; key.pat = x.pat;
  key.ws_generation = yh_workspace_generation();
  // the kernel-selection switches the launchers read from the environment (A/B hooks, tests) are part of what a
  // captured graph froze: a cached graph must not outlive a change of any of them
  unsigned long long h = 1469598103934665603ull;   // FNV-1a over "NAME=value;"
  for (const char *name : {"YH_RD_PATH", "YH_TILE_RK", "YH_EULER_KERNEL", "YH_RK_KERNEL", "YH_SOLID_RK", "YH_EULER_FEED",
                           "YH_MARCH_TILING", "YH_MARCH_R", "YH_FAST_W", "YH_RK_W"}) {
    const char *v = getenv(name);
    for (const char *c = v ? v : ""; *c; c++) h = (h ^ (unsigned char)*c) * 1099511628211ull;
    h = (h ^ 0x3bu) * 1099511628211ull;
  }
  key.arith = (long long)yh_arithmetic() | (long long)((h >> 8) << 8);
  return key;
}

static GraphEntry *graph_find(const GraphKey &key, int dev) {
  for (int q = 0; q < YH_GRAPH_CACHE; q++) {
    GraphEntry &g = g_graphs[q];
    if (g.exec && g.dev == dev && memcmp(&g.key, &key, sizeof(key)) == 0) return &g;
  }
  return nullptr;
}

// Capture one chunk (state in (cu,cv) before and after) into the least recently used cache slot.
static int graph_capture(const AdvanceCtx &x, const GraphKey &key, int dev, double *cu, double *cv,
                         double *nu, double *nv, cudaStream_t gs, GraphEntry **out) {
  GraphEntry *victim = &g_graphs[0];
  for (int q = 0; q < YH_GRAPH_CACHE; q++) {
    if (!g_graphs[q].exec) { victim = &g_graphs[q]; break; }
    if (g_graphs[q].stamp < victim->stamp) victim = &g_graphs[q];
  }
  if (victim->exec) { cudaGraphExecDestroy(victim->exec); victim->exec = nullptr; }
  cudaGraph_t graph = nullptr;
  YH_CUDA(cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal));
  int canon = 0, inB = 0;
  int rc = advance_plain(x, YH_GRAPH_CHUNK, canon, cu, cv, nu, nv, inB, gs);
  cudaError_t ce = cudaStreamEndCapture(gs, &graph);
  if (rc != YH_OK || ce != cudaSuccess || inB != 0) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    if (rc == YH_OK) { yh_set_error("graph capture of the step loop failed"); rc = YH_ERR_CUDA; }
    return rc;
  }
  cudaGraphExec_t exec = nullptr;
  ce = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) { cudaGetLastError(); yh_set_error("cudaGraphInstantiate failed"); return YH_ERR_CUDA; }
  victim->key = key; victim->exec = exec; victim->dev = dev;
  *out = victim;
  return YH_OK;
}

void yh_graphs_release(void) {
  for (int q = 0; q < YH_GRAPH_CACHE; q++)
    if (g_graphs[q].exec) { cudaGraphExecDestroy(g_graphs[q].exec); g_graphs[q].exec = nullptr; }
}

// The whole-sheet step loop used by yh_rd_advance and the headless driver.  Order of passes:
// [one plain chunk when the input is raw or the graph is new] [graph replays] [plain remainder],
// so the LAST pass is the shortest one, as in the plain loop (yh_sim_tips relies on a final
// single-step pass).
int yh_advance_whole(const yh_params *p, const YhK &k, int nsteps, int tb, int canon_in, double *uA,
                     double *vA, double *uB, double *vB, const uint8_t *solid, const uint8_t *pat,
                     int row0, int row1, int *result_in_B, int *last_T, cudaStream_t st) {
  AdvanceCtx x{p, k, tb, row0, row1, solid, pat};
  double *cu = uA, *cv = vA, *nu = uB, *nv = vB;
  int inB = 0, canon = canon_in, rc, dev = 0;
  if (last_T) {   // length of the final pass
    const bool fast = pat || yh_rd_fast_supported(k, 1);
    int T = fast ? tb : 1, left = nsteps;
    while (fast && left > 0 && left % T) T >>= 1;   // tb is a power of two: the tail runs T/2, T/4, ...
    *last_T = nsteps > 0 ? T : 0;
  }
  const bool whole = p->jg0 == 0 && p->ny_global == p->ny && row0 == 0 && row1 == p->ny;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) cudaGetLastError();
  int nch = nsteps / YH_GRAPH_CHUNK;
  YH_CUDA(cudaGetDevice(&dev));
  if (whole && !yh_preload_only && cap == cudaStreamCaptureStatusNone && nch >= 3 && dev < YH_MAX_DEV &&
      yh_graphs_enabled((long long)p->nx * p->ny)) {
    cudaStream_t gs;
    rc = graph_stream(dev, &gs);
    if (rc != YH_OK) return rc;
    GraphKey key = graph_key(x, cu, cv, nu, nv);
    GraphEntry *e = graph_find(key, dev);
    if (canon || !e) {   // raw input, or first use of these kernel variants: one chunk of plain launches
      rc = advance_plain(x, YH_GRAPH_CHUNK, canon, cu, cv, nu, nv, inB, st);
      if (rc != YH_OK) return rc;
      nch--;
    }
    if (!e) rc = graph_capture(x, key, dev, cu, cv, nu, nv, gs, &e);
    if (rc == YH_OK) {
      e->stamp = ++g_graph_clock;
      YH_CUDA(cudaEventRecord(g_graph_ev[dev][0], st));
      YH_CUDA(cudaStreamWaitEvent(gs, g_graph_ev[dev][0], 0));
      for (int q = 0; q < nch; q++) YH_CUDA(cudaGraphLaunch(e->exec, gs));
      YH_CUDA(cudaEventRecord(g_graph_ev[dev][1], gs));
      YH_CUDA(cudaStreamWaitEvent(st, g_graph_ev[dev][1], 0));
      nch = 0;
    } else if (rc != YH_ERR_CUDA) {
      return rc;
    }   // capture refused: the rest of the run goes through plain launches
    rc = advance_plain(x, nch * YH_GRAPH_CHUNK + nsteps % YH_GRAPH_CHUNK, canon, cu, cv, nu, nv, inB, st);
    if (rc != YH_OK) return rc;
    *result_in_B = inB;
    return YH_OK;
  }
  rc = advance_plain(x, nsteps, canon, cu, cv, nu, nv, inB, st);
  if (rc != YH_OK) return rc;
  *result_in_B = inB;
  return YH_OK;
}

extern "C" {

int yh_rd_advance(const yh_params *p, int nsteps, int tb_steps, int flags, double *uA, double *vA,
                  double *uB, double *vB, const uint8_t *solid, int stim_mouse, int point_x,
                  int point_y, int row0, int row1, int *result_in_B, void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  YH_REQUIRE(nsteps >= 0 && result_in_B, "bad nsteps / result_in_B");
  YH_REQUIRE(tb_steps >= 0 && tb_steps <= 4 && tb_steps != 3, "tb_steps must be 0 (auto), 1, 2 or 4");
  rc = validate_rd(p, uA, vA, uB, vB, solid, row0, row1,
                   nsteps * (p ? p->timeIntOrder * yh_rd_radius(p) : 1));
  if (rc != YH_OK) return rc;
  YhK k = yh_make_k(p);
  k.stim = stim_mouse != 0; k.px = point_x; k.py = point_y;
  cudaStream_t st = (cudaStream_t)stream;
  const int tb = tb_steps ? tb_steps : 4;
  const int canon = (flags & YH_RD_INPUT_CANONICAL) ? 0 : 1;   // only the first pass can see user data
  // obstacle masks + Euler: the temporally blocked kernel reads per-cell mask patterns, made
  // once per call (mask contents belong to the caller and may change between calls)
  uint8_t *pat = nullptr;
  if (flags & YH_RD_SOLID_IS_PATTERNS) {
    if (!yh_rd_fast_solid_supported(k, 1)) {
      yh_set_error("YH_RD_SOLID_IS_PATTERNS needs the masked Euler / no-flux mode");
      return YH_ERR_UNSUPPORTED;
    }
    pat = const_cast<uint8_t *>(solid);
  } else if (yh_rd_fast_solid_supported(k, 1) && nsteps > 1) {
    rc = yh_workspace((size_t)p->nx * p->ny, (void **)&pat, 4);
    if (rc != YH_OK) return rc;
    rc = yh_rd_solid_patterns(k, solid, pat, st);
    if (rc != YH_OK) return rc;
  }
  return yh_advance_whole(p, k, nsteps, tb, canon, uA, vA, uB, vB, solid, pat, row0, row1, result_in_B, nullptr, st);
}

int yh_rd_mask_patterns(const yh_params *p, const uint8_t *solid, uint8_t *patterns, void *stream) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  YH_REQUIRE(p && solid && patterns, "null pointer");
  YH_REQUIRE(p->nx >= 4 && p->ny >= 1 && p->jg0 >= 0 && p->jg0 + p->ny <= p->ny_global, "bad slab");
  return yh_rd_solid_patterns(yh_make_k(p), solid, patterns, (cudaStream_t)stream);
}

// solve_matrix, symmetryReduction.cu:386-416 (host; same operation order; no leak)
int yh_solve_matrix(const double c_in[3], const double phi[3], const double Int[12],
                    double c_out[3]) {
  (void)c_in;
  if (!phi || !Int || !c_out) { yh_set_error("yh_solve_matrix: null pointer"); return YH_ERR_INVALID_ARG; }
  const double pt = phi[2];
  yh_solve3(Int, cos(pt), sin(pt), c_out);
  return YH_OK;
}

}  // extern "C"
