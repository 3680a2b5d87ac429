// sim.cu -- headless driver: the display() step loop of the reference (main.cu:862-1043)
// without GLUT/OpenGL, for one sheet or a batch of independent sheets (parameter sweeps).
// State lives in HBM for the whole run; the per-step blocking 16-byte D2H of
// singleCell_wrapper (singleCell.cu:28) becomes a device-side trace buffer drained once.
#include <string.h>
#include <vector>

#include "yh_common.cuh"

struct yh_sim {
  yh_params p;
  int n_sims, device;
  size_t n;              // cells per sheet
  double *u[2], *v[2];
  int cur;
  uint8_t *solid;
  double *trace_d;
  size_t trace_cap;
  int *tip_count_d;
  yh_tip *tip_vec_d;
  int *period_d;
  std::vector<int> period_h;
  int duration_it;
  int px, py;
  int count;
  int have_prev;         // other buffer holds the state exactly one step back
  int raw_input;         // current state came from the host and may hold -0.0
  double *vt[2], *adv[2];   // velTan and advection field (symmetry reduction), lazily allocated
  double c[3], phi[3];   // drift velocities and frame phase (main.cu:60-61)
  double *apd[6];        // APD1 APD2 sAPD dAPD back front, n_sims sheets each (lazily allocated)
  uint8_t *apd_first, *stim_area;
  double *sr_d, *sr_log_d;   // device-resident (c, phi, cos, sin) and its per-step record (yh_sim_run_sr_device)
  size_t sr_log_cap;
  unsigned long long *probe_slot_d;   // sample slot of the graph-replayed probe (yh_sim_run with a trace)
  uint8_t *pat;          // mask patterns of the temporally blocked Euler kernel (yh_rd_solid_patterns)
  int apd_init;          // sAPD / dAPD hold values for every cell (one full pass done)
  int have_stim_area;    // stim_area holds a mask uploaded by an earlier yh_sim_run_apd call
  cudaStream_t st;
};

namespace {

__global__ void probe_batch_kernel(const double *u, const double *v, double *out, long long idx,
                                   long long stride, int nsims) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsims) return;
  out[2 * s] = u[idx + stride * s];
  out[2 * s + 1] = v[idx + stride * s];
}

// Same probe for graph-replayed loops: the sample slot comes from a device-resident counter, so the
// launch arguments do not change from step to step (one block: nsims <= blockDim.x).
__global__ void probe_counted_kernel(const double *u, const double *v, double *out, long long idx,
                                     long long stride, int nsims, unsigned long long *slot) {
  __shared__ unsigned long long s_slot;
  if (threadIdx.x == 0) s_slot = *slot;
  __syncthreads();
  const int z = threadIdx.x;
  if (z < nsims) {
    out[2 * (s_slot * nsims + z)] = u[idx + stride * z];
    out[2 * (s_slot * nsims + z) + 1] = v[idx + stride * z];
  }
  if (threadIdx.x == 0) *slot = s_slot + 1;
}

struct DevGuard {
  int prev;
  explicit DevGuard(int d) { cudaGetDevice(&prev); cudaSetDevice(d); }
  ~DevGuard() { cudaSetDevice(prev); }
};

}  // namespace

extern "C" {

int yh_sim_create(yh_sim **out, const yh_params *p, int n_sims, int device) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  YH_REQUIRE(out && p && n_sims >= 1, "bad arguments");
  YH_REQUIRE(p->jg0 == 0 && p->ny_global == p->ny, "yh_sim drives whole sheets");
  YH_REQUIRE(device >= 0 && device < yh_device_count(), "no such device");
  DevGuard g(device);
  yh_sim *s = new yh_sim();
  memset(&s->p, 0, sizeof(s->p));
  s->p = *p;
  s->n_sims = n_sims; s->device = device;
  s->n = (size_t)p->nx * p->ny;
  s->cur = 0; s->solid = nullptr; s->pat = nullptr; s->sr_d = nullptr; s->sr_log_d = nullptr; s->sr_log_cap = 0; s->probe_slot_d = nullptr; s->trace_d = nullptr; s->trace_cap = 0;
  s->period_d = nullptr; s->duration_it = 0; s->count = 0; s->have_prev = 0; s->raw_input = 1;
  s->px = p->nx / 2; s->py = p->ny / 2;   // param.point, saveFiles.cu:180
  s->vt[0] = s->vt[1] = s->adv[0] = s->adv[1] = nullptr;
  for (int q = 0; q < 6; q++) s->apd[q] = nullptr;
  s->apd_first = s->stim_area = nullptr;
  s->apd_init = 0;
  for (int q = 0; q < 3; q++) { s->c[q] = 0.0; s->phi[q] = 0.0; }
  const size_t bytes = s->n * n_sims * sizeof(double);
  for (int b = 0; b < 2; b++) {
    YH_CUDA(cudaMalloc(&s->u[b], bytes));
    YH_CUDA(cudaMalloc(&s->v[b], bytes));
    YH_CUDA(cudaMemset(s->u[b], 0, bytes));
    YH_CUDA(cudaMemset(s->v[b], 0, bytes));
  }
  YH_CUDA(cudaMalloc(&s->tip_count_d, sizeof(int)));
  YH_CUDA(cudaMemset(s->tip_count_d, 0, sizeof(int)));
  YH_CUDA(cudaMalloc(&s->tip_vec_d, sizeof(yh_tip) * (size_t)YH_TIPVECSIZE));
  YH_CUDA(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
  *out = s;
  return YH_OK;
}

int yh_sim_destroy(yh_sim *s) {
  if (!s) return YH_OK;
  DevGuard g(s->device);
  cudaStreamSynchronize(s->st);
  for (int b = 0; b < 2; b++) { cudaFree(s->u[b]); cudaFree(s->v[b]); }
  cudaFree(s->solid); cudaFree(s->pat); cudaFree(s->sr_d); cudaFree(s->sr_log_d); cudaFree(s->probe_slot_d); cudaFree(s->trace_d); cudaFree(s->tip_count_d); cudaFree(s->tip_vec_d);
  cudaFree(s->period_d);
  cudaFree(s->vt[0]); cudaFree(s->vt[1]); cudaFree(s->adv[0]); cudaFree(s->adv[1]);
  for (int q = 0; q < 6; q++) cudaFree(s->apd[q]);
  cudaFree(s->apd_first); cudaFree(s->stim_area);
  cudaStreamDestroy(s->st);
  delete s;
  return YH_OK;
}

int yh_sim_set_state(yh_sim *s, const double *u_h, const double *v_h) {
  YH_REQUIRE(s && u_h && v_h, "null pointer");
  DevGuard g(s->device);
  const size_t bytes = s->n * s->n_sims * sizeof(double);
  YH_CUDA(cudaMemcpyAsync(s->u[s->cur], u_h, bytes, cudaMemcpyHostToDevice, s->st));
  YH_CUDA(cudaMemcpyAsync(s->v[s->cur], v_h, bytes, cudaMemcpyHostToDevice, s->st));
  YH_CUDA(cudaStreamSynchronize(s->st));
  s->have_prev = 0;
  s->raw_input = 1;
  return YH_OK;
}

int yh_sim_get_state(yh_sim *s, double *u_h, double *v_h) {
  YH_REQUIRE(s && u_h && v_h, "null pointer");
  DevGuard g(s->device);
  const size_t bytes = s->n * s->n_sims * sizeof(double);
  YH_CUDA(cudaMemcpyAsync(u_h, s->u[s->cur], bytes, cudaMemcpyDeviceToHost, s->st));
  YH_CUDA(cudaMemcpyAsync(v_h, s->v[s->cur], bytes, cudaMemcpyDeviceToHost, s->st));
  YH_CUDA(cudaStreamSynchronize(s->st));
  return YH_OK;
}

int yh_sim_set_solid(yh_sim *s, const uint8_t *solid_h) {
  YH_REQUIRE(s && solid_h, "null pointer");
  DevGuard g(s->device);
  if (!s->solid) YH_CUDA(cudaMalloc(&s->solid, s->n));
  if (!s->pat) YH_CUDA(cudaMalloc(&s->pat, s->n));
  YH_CUDA(cudaMemcpy(s->solid, solid_h, s->n, cudaMemcpyHostToDevice));
  int rc = yh_rd_solid_patterns(yh_make_k(&s->p), s->solid, s->pat, s->st);
  if (rc != YH_OK) return rc;
  YH_CUDA(cudaStreamSynchronize(s->st));
  return YH_OK;
}

// initGates, main.cu:606-618: u = 1 for i < nx/8, v = 1 for j >= ny/2 (every sheet)
int yh_sim_cross_field_ic(yh_sim *s) {
  YH_REQUIRE(s != nullptr, "null sim");
  const int nx = s->p.nx, ny = s->p.ny;
  std::vector<double> u(s->n * s->n_sims, 0.0), v(s->n * s->n_sims, 0.0);
  for (int z = 0; z < s->n_sims; z++) {
    double *uz = u.data() + s->n * z, *vz = v.data() + s->n * z;
    for (int j = 0; j < ny; j++)
      for (int i = 0; i < nx / 8; i++) uz[i + (size_t)nx * j] = 1.0;
    for (int j = ny / 2; j < ny; j++)
      for (int i = 0; i < nx; i++) vz[i + (size_t)nx * j] = 1.0;
  }
  s->count = 0;
  return yh_sim_set_state(s, u.data(), v.data());
}

int yh_sim_set_point(yh_sim *s, int x, int y) {
  YH_REQUIRE(s && x >= 0 && x < s->p.nx && y >= 0 && y < s->p.ny, "bad point");
  s->px = x; s->py = y;
  return YH_OK;
}

int yh_sim_set_pacing(yh_sim *s, const int *period_it, int duration_it) {
  YH_REQUIRE(s != nullptr, "null sim");
  DevGuard g(s->device);
  if (!period_it) {   // pacing off
    cudaFree(s->period_d); s->period_d = nullptr; s->period_h.clear();
    return YH_OK;
  }
  YH_REQUIRE(duration_it >= 0, "negative duration");
  s->period_h.assign(period_it, period_it + s->n_sims);
  if (!s->period_d) YH_CUDA(cudaMalloc(&s->period_d, sizeof(int) * s->n_sims));
  YH_CUDA(cudaMemcpy(s->period_d, period_it, sizeof(int) * s->n_sims, cudaMemcpyHostToDevice));
  s->duration_it = duration_it;
  return YH_OK;
}

static int sim_run_impl(yh_sim *s, int nsteps, int tb_steps, double *trace_h, bool sync);

// nsteps x { reactionDiffusion ; swap ; probe }  (main.cu:869-885, 1040)
int yh_sim_run(yh_sim *s, int nsteps, int tb_steps, double *trace_h) {
  return sim_run_impl(s, nsteps, tb_steps, trace_h, true);
}

static int sim_run_impl(yh_sim *s, int nsteps, int tb_steps, double *trace_h, bool sync) {
  YH_REQUIRE(s && nsteps >= 0, "bad arguments");
  YH_REQUIRE(tb_steps >= 0 && tb_steps <= 4 && tb_steps != 3, "tb_steps must be 0 (auto), 1, 2 or 4");
  DevGuard g(s->device);
  YhK k = yh_make_k(&s->p);
  k.px = s->px; k.py = s->py;
  const bool pacing = s->period_d != nullptr;
  const long long stride = (long long)s->n;
  int tb = tb_steps ? tb_steps : 4;
  if (trace_h) {
    tb = 1;   // one probe sample per step, taken BEFORE the step (one-step lag of main.cu:1040)
    const size_t need = (size_t)2 * nsteps * s->n_sims;
    if (s->trace_cap < need) {
      cudaFree(s->trace_d); s->trace_d = nullptr; s->trace_cap = 0;
      YH_CUDA(cudaMalloc(&s->trace_d, need * sizeof(double)));
      s->trace_cap = need;
    }
  }
  const uint8_t *pat = (s->pat && yh_rd_fast_solid_supported(k, 1)) ? s->pat : nullptr;
  const bool fast1 = pat != nullptr || yh_rd_fast_supported(k, 1) != 0;
  if (!trace_h && !pacing && s->n_sims == 1 && nsteps > 0) {
    // a single unpaced sheet: the library's step loop (CUDA-graph replay when the sheet is small)
    int inB = 0, last_T = 0;
    const int c = s->cur, o = c ^ 1;
    int rc = yh_advance_whole(&s->p, k, nsteps, tb, s->raw_input, s->u[c], s->v[c], s->u[o], s->v[o],
                              s->solid, pat, 0, s->p.ny, &inB, &last_T, s->st);
    if (rc != YH_OK) return rc;
    if (inB) s->cur = o;
    s->raw_input = 0;
    s->have_prev = (last_T == 1);
    s->count += nsteps;
    if (sync) YH_CUDA(cudaStreamSynchronize(s->st));
    return YH_OK;
  }
  const bool tile = !pat && yh_rd_prefer_tile((long long)s->n * s->n_sims) != 0;
  // one RD launch (all sheets) of T time steps from buffers c to o
  auto rd_launch = [&](int c, int o, int T) -> int {
    if (fast1 && tile)
      return yh_launch_rd_tile_euler(k, T, s->u[c], s->v[c], s->u[o], s->v[o], s->n_sims, stride,
                                     s->period_d, s->duration_it, s->count, s->st);
    if (fast1)
      return yh_launch_rd_fast_paced(k, T, s->u[c], s->v[c], s->u[o], s->v[o], s->n_sims, stride,
                                     s->period_d, s->duration_it, s->count, s->raw_input, s->st, nullptr, pat);
    int rc = YH_OK;
    for (int z = 0; z < s->n_sims && rc == YH_OK; z++) {
      YhK kz = k;
      if (pacing) {
        const int per = s->period_h[z];
        kz.stim = per > 0 && (s->count % per) <= s->duration_it;
      }
      if (tile && yh_rd_tile_rk_supported(kz))
        rc = yh_launch_rd_tile_rk(kz, s->u[c] + s->n * z, s->v[c] + s->n * z, s->u[o] + s->n * z,
                                  s->v[o] + s->n * z, nullptr, nullptr, s->st);
      else if (yh_rd_rk_supported(kz))
        rc = yh_launch_rd_rk(kz, s->u[c] + s->n * z, s->v[c] + s->n * z, s->u[o] + s->n * z,
                             s->v[o] + s->n * z, nullptr, nullptr, s->solid, s->st);
      else
        rc = yh_launch_rd_generic(kz, s->u[c] + s->n * z, s->v[c] + s->n * z, s->u[o] + s->n * z,
                                  s->v[o] + s->n * z, nullptr, nullptr, s->solid, s->st);
    }
    return rc;
  };
  int left = nsteps, step = 0;
  // The reference's own loop {step; swap; probe} (main.cu:879-885, 1040) on a small sheet: chunks of
  // 64 {probe, step} pairs replayed from a CUDA graph (the probe's sample slot is a device-resident
  // counter, so no launch argument changes); head (raw input) and tail run as plain launches.
  constexpr int CH = 64;
  if (trace_h && !pacing && s->n_sims <= 128 && nsteps >= 3 * CH &&
      yh_graphs_enabled((long long)s->n * s->n_sims)) {
    if (!s->probe_slot_d) YH_CUDA(cudaMalloc(&s->probe_slot_d, sizeof(unsigned long long)));
    YH_CUDA(cudaMemsetAsync(s->probe_slot_d, 0, sizeof(unsigned long long), s->st));
    const long long pidx = (long long)s->px + (long long)s->p.nx * s->py;
    // Euler on tiles: the probe is fused into the step kernel (the owner thread of the electrode
    // cell records it at every time level), four steps per launch; otherwise {probe, step} pairs
    const bool fused = fast1 && tile;
    auto chunk = [&](int n) -> int {
      for (int q = 0; q < n;) {
        const int c = s->cur, o = c ^ 1;
        int T = 1;
        if (fused) {
          T = 4;
          while (T > n - q) T >>= 1;
          int rc = yh_launch_rd_tile_euler(k, T, s->u[c], s->v[c], s->u[o], s->v[o], s->n_sims, stride,
                                           nullptr, 0, s->count, s->st, s->trace_d, s->probe_slot_d);
          if (rc != YH_OK) return rc;
          rc = yh_slot_bump(s->probe_slot_d, T, s->st);
          if (rc != YH_OK) return rc;
        } else {
          probe_counted_kernel<<<1, 128, 0, s->st>>>(s->u[c], s->v[c], s->trace_d, pidx, stride, s->n_sims,
                                                     s->probe_slot_d);
          YH_LAUNCH_CHECK();
          int rc = rd_launch(c, o, 1);
          if (rc != YH_OK) return rc;
        }
        s->cur = o; s->raw_input = 0; s->count += T; s->have_prev = (T == 1);
        q += T;
      }
      return YH_OK;
    };
    int rc = chunk(CH);   // raw input and first use of the kernel variants: plain launches
    if (rc != YH_OK) return rc;
    left -= CH;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    const int count_before = s->count;
    YH_CUDA(cudaStreamBeginCapture(s->st, cudaStreamCaptureModeThreadLocal));
    rc = chunk(CH);
    cudaError_t ce = cudaStreamEndCapture(s->st, &graph);
    s->count = count_before;   // the capture launched nothing
    if (rc == YH_OK && ce == cudaSuccess && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
      for (; left >= CH; left -= CH) { YH_CUDA(cudaGraphLaunch(exec, s->st)); s->count += CH; }
    } else {
      cudaGetLastError();
      if (rc != YH_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    }
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    rc = chunk(left);
    if (rc != YH_OK) return rc;
    left = 0;
  }
  while (left > 0) {
    int T = 1;
    if (fast1) { T = tb; while (T > left) T >>= 1; }
    const int c = s->cur, o = c ^ 1;
    if (trace_h) {
      probe_batch_kernel<<<(s->n_sims + 127) / 128, 128, 0, s->st>>>(
          s->u[c], s->v[c], s->trace_d + (size_t)2 * step * s->n_sims,
          (long long)s->px + (long long)s->p.nx * s->py, stride, s->n_sims);
      YH_LAUNCH_CHECK();
    }
    int rc = rd_launch(c, o, T);
    if (rc != YH_OK) return rc;
    s->cur = o;
    s->raw_input = 0;
    s->have_prev = (T == 1);
    s->count += T;
    left -= T; step += T;
  }
  if (trace_h && nsteps > 0) {
    YH_CUDA(cudaMemcpyAsync(trace_h, s->trace_d, sizeof(double) * 2 * nsteps * s->n_sims,
                            cudaMemcpyDeviceToHost, s->st));
  }
  if (sync) YH_CUDA(cudaStreamSynchronize(s->st));
  return YH_OK;
}

int yh_sim_run_host(yh_sim *s, const double *u_in_h, const double *v_in_h, double *u_out_h,
                    double *v_out_h, int nsteps, int tb_steps) {
  YH_REQUIRE(s && u_in_h && v_in_h && u_out_h && v_out_h, "null pointer");
  DevGuard g(s->device);
  const size_t bytes = s->n * s->n_sims * sizeof(double);
  YH_CUDA(cudaMemcpyAsync(s->u[s->cur], u_in_h, bytes, cudaMemcpyHostToDevice, s->st));
  YH_CUDA(cudaMemcpyAsync(s->v[s->cur], v_in_h, bytes, cudaMemcpyHostToDevice, s->st));
  s->raw_input = 1;
  int rc = yh_sim_run(s, nsteps, tb_steps, nullptr);
  if (rc != YH_OK) return rc;
  YH_CUDA(cudaMemcpyAsync(u_out_h, s->u[s->cur], bytes, cudaMemcpyDeviceToHost, s->st));
  YH_CUDA(cudaMemcpyAsync(v_out_h, s->v[s->cur], bytes, cudaMemcpyDeviceToHost, s->st));
  YH_CUDA(cudaStreamSynchronize(s->st));
  return YH_OK;
}

// Tips between the last two states (tip_wrapper as called at main.cu:963: present = newest).
// Sheet 0 only.  Requires the last pass to have been a single step (run with tb_steps = 1, or
// any run whose final pass has T = 1).
int yh_sim_tips(yh_sim *s, yh_tip *tips_h, int capacity, int *count_out) {
  YH_REQUIRE(s && tips_h && count_out && capacity >= 0, "bad arguments");
  if (!s->have_prev) {
    yh_set_error("yh_sim_tips: previous state not held (end the run with a single-step pass)");
    return YH_ERR_UNSUPPORTED;
  }
  DevGuard g(s->device);
  const double t = s->p.dt * (double)s->count;   // param.physicalTime, main.cu:885
  int rc = yh_tip_track(&s->p, s->u[s->cur ^ 1], s->u[s->cur], nullptr, s->tip_count_d,
                        s->tip_vec_d, YH_TIPVECSIZE, t, s->p.tipAlgorithm, s->st);
  if (rc != YH_OK) return rc;
  int n = 0;
  YH_CUDA(cudaMemcpyAsync(&n, s->tip_count_d, sizeof(int), cudaMemcpyDeviceToHost, s->st));
  YH_CUDA(cudaStreamSynchronize(s->st));
  *count_out = n;
  if (n > YH_TIPVECSIZE) { yh_set_error("tip list overflow: %d > %d", n, YH_TIPVECSIZE); return YH_ERR_CAPACITY; }
  const int m = n < capacity ? n : capacity;
  if (m > 0) YH_CUDA(cudaMemcpy(tips_h, s->tip_vec_d, sizeof(yh_tip) * (size_t)m, cudaMemcpyDeviceToHost));
  return YH_OK;
}

static int sr_alloc(yh_sim *s) {   // velTan and the advection field, zeroed (main.cu:431-434)
  const size_t bytes = s->n * sizeof(double);
  if (s->vt[0]) return YH_OK;
  for (int q = 0; q < 2; q++) {
    YH_CUDA(cudaMalloc(&s->vt[q], bytes));
    YH_CUDA(cudaMalloc(&s->adv[q], bytes));
    YH_CUDA(cudaMemsetAsync(s->vt[q], 0, bytes, s->st));
    YH_CUDA(cudaMemsetAsync(s->adv[q], 0, bytes, s->st));
  }
  return YH_OK;
}

// One symmetry-reduction step per iteration, as display() does when param.reduceSym
// (main.cu:894-954), with slice+trapz fused and Cxy fused into the advection.
int yh_sim_run_sr(yh_sim *s, int nsteps, double *c_phi_h) {
  YH_REQUIRE(s && nsteps >= 0, "bad arguments");
  YH_REQUIRE(s->n_sims == 1, "symmetry reduction drives a single sheet");
  DevGuard g(s->device);
  int rc0 = sr_alloc(s);
  if (rc0 != YH_OK) return rc0;
  const yh_params *p = &s->p;
  double I[12];
  for (int it = 0; it < nsteps; it++) {
    const int c = s->cur, o = c ^ 1;
    // u* = RD(u^n), velTan = rhs/dt                                       (main.cu:896)
    int rc = yh_rd_step(p, s->u[c], s->v[c], s->u[o], s->v[o], s->vt[0], s->vt[1], s->solid, 0, s->px,
                        s->py, 0, p->ny, s->st);
    if (rc != YH_OK) return rc;
    // tips between u^n (present) and u* (past), every step                 (main.cu:900)
    rc = yh_tip_track(p, s->u[o], s->u[c], nullptr, s->tip_count_d, s->tip_vec_d, YH_TIPVECSIZE,
                      p->dt * (double)s->count, p->tipAlgorithm, s->st);
    if (rc != YH_OK) return rc;
    if (c_phi_h) {                                                       // main.cu:902-903
      for (int q = 0; q < 3; q++) { c_phi_h[6 * it + q] = s->c[q]; c_phi_h[6 * it + 3 + q] = s->phi[q]; }
    }
    // phase-condition integrals of the tangent fields of u^n              (main.cu:906-923)
    rc = yh_sr_integrals(p, s->u[c], s->v[c], s->vt[0], s->vt[1], s->adv[0], s->adv[1], I, s->tip_count_d,
                         s->tip_vec_d, s->count, s->st);
    if (rc != YH_OK) return rc;
    if (s->count == 0) {   // first step: solve, rebuild the frame velocity, slice again (main.cu:910-921)
      yh_solve_matrix(s->c, s->phi, I, s->c);
      rc = yh_cxy_field(p, s->adv[0], s->adv[1], s->c, s->phi, s->solid, s->st);
      if (rc != YH_OK) return rc;
      rc = yh_sr_integrals(p, s->u[c], s->v[c], s->vt[0], s->vt[1], s->adv[0], s->adv[1], I,
                           s->tip_count_d, s->tip_vec_d, s->count, s->st);
      if (rc != YH_OK) return rc;
    }
    yh_solve_matrix(s->c, s->phi, I, s->c);                               // main.cu:926
    // u^{n+1} = BFECC(u*) in the frame moving with (c, phi); result lands in the `c` buffers,
    // exactly the reference's ping-pong without swap                     (main.cu:930-932)
    rc = yh_advect_bfecc_cphi(p, s->u[o], s->v[o], s->u[c], s->v[c], s->c, s->phi, s->adv[0], s->adv[1],
                              s->solid, s->st);
    if (rc != YH_OK) return rc;
    for (int q = 0; q < 3; q++) s->phi[q] = s->phi[q] + s->c[q] * p->dt;   // main.cu:936-938 (sticky start)
    s->count++;
    s->have_prev = 0;
    s->raw_input = 0;
  }
  YH_CUDA(cudaStreamSynchronize(s->st));
  return YH_OK;
}

// The same symmetry-reduction steps with the drift solve RESIDENT ON THE DEVICE: the 12 integrals,
// the 3x3 solve, the frame update and the (c, phi) record never leave the GPU, so a step is five
// back-to-back launches {RD, tips, integral rows, integrals+solve, BFECC} with no host round trip
// (the reference blocks the host 12 times per step, integralTrapz.cu:97-179).  cos/sin(phi.t) come
// from libdevice instead of libm: (c, phi) agree with yh_sim_run_sr to rounding, not bit for bit.
// The very first step of a run (count == 0: two solves, main.cu:910-921) goes through the host path.
int yh_sim_run_sr_device(yh_sim *s, int nsteps, double *c_phi_h) {
  YH_REQUIRE(s && nsteps >= 0, "bad arguments");
  YH_REQUIRE(s->n_sims == 1, "symmetry reduction drives a single sheet");
  int done = 0;
  if (s->count == 0 && nsteps > 0) {
    int rc = yh_sim_run_sr(s, 1, c_phi_h);
    if (rc != YH_OK) return rc;
    done = 1;
  }
  if (done == nsteps) return YH_OK;
  DevGuard g(s->device);
  int rc0 = sr_alloc(s);
  if (rc0 != YH_OK) return rc0;
  const yh_params *p = &s->p;
  const int n = nsteps - done;
  if (!s->sr_d) YH_CUDA(cudaMalloc(&s->sr_d, YH_SR_WORDS * sizeof(double)));
  if (s->sr_log_cap < (size_t)n) {
    cudaFree(s->sr_log_d); s->sr_log_d = nullptr; s->sr_log_cap = 0;
    YH_CUDA(cudaMalloc(&s->sr_log_d, (size_t)6 * n * sizeof(double)));
    s->sr_log_cap = (size_t)n;
  }
  // upload (c, phi, step): the closing kernel of a step first applies the deferred phi += c*dt of
  // the step before it -- except for the first step of this run, where the host has applied it
  double h[YH_SR_WORDS] = {0};
  for (int q = 0; q < 3; q++) { h[YH_SR_C + q] = s->c[q]; h[YH_SR_PHI + q] = s->phi[q]; }
  h[YH_SR_STEP] = (double)s->count;
  YH_CUDA(cudaMemcpyAsync(s->sr_d, h, sizeof(h), cudaMemcpyHostToDevice, s->st));
  YH_CUDA(cudaStreamSynchronize(s->st));   // h lives on this stack frame
  const double step0 = (double)s->count;
  const int c = s->cur, o = c ^ 1;          // RD c -> o, BFECC o -> c: the same buffers every step
  // One step = {RD, tips, integral rows + closing, BFECC}.  No argument changes from step to step
  // (the step counter, the time tag of the tips and the record slot live on the device), so a
  // chunk of steps is captured once and replayed as a CUDA graph; the tail runs as plain launches.
  auto one_step = [&](cudaStream_t st) -> int {
    int rc = yh_rd_step(p, s->u[c], s->v[c], s->u[o], s->v[o], s->vt[0], s->vt[1], s->solid, 0, s->px,
                        s->py, 0, p->ny, st);
    if (rc != YH_OK) return rc;
    // tips between u^n (present) and u* (past); tagged t = dt*count (main.cu:900)
    rc = yh_tip_track_device_step(p, s->u[o], s->u[c], s->tip_count_d, s->tip_vec_d, YH_TIPVECSIZE,
                                  s->sr_d, 0, st);
    if (rc != YH_OK) return rc;
    rc = yh_sr_integrals_solve_device(p, s->u[c], s->v[c], s->vt[0], s->vt[1], s->adv[0], s->adv[1],
                                      s->tip_count_d, s->tip_vec_d, s->sr_d, s->sr_log_d, step0, st);
    if (rc != YH_OK) return rc;
    return yh_advect_bfecc_device_c(p, s->u[o], s->v[o], s->u[c], s->v[c], s->sr_d, s->adv[0], s->adv[1],
                                    s->solid, st);
  };
  constexpr int G = 8;
  int left = n;
  if (n >= 4 * G && yh_graphs_enabled((long long)s->n)) {
    int rc = one_step(s->st);              // first use of every kernel variant outside the capture
    if (rc != YH_OK) return rc;
    left--;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    YH_CUDA(cudaStreamBeginCapture(s->st, cudaStreamCaptureModeThreadLocal));
    for (int g = 0; g < G && rc == YH_OK; g++) rc = one_step(s->st);
    cudaError_t ce = cudaStreamEndCapture(s->st, &graph);
    if (rc == YH_OK && ce == cudaSuccess && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
      for (; left >= G; left -= G) YH_CUDA(cudaGraphLaunch(exec, s->st));
    } else {
      cudaGetLastError();                  // capture refused: plain launches do the whole run
      if (rc != YH_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    }
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
  }
  for (; left > 0; left--) {
    int rc = one_step(s->st);
    if (rc != YH_OK) return rc;
  }
  s->count += n;
  int rc = yh_sr_flush_phi_device(s->sr_d, p->dt, s->st);   // the last step's phi += c*dt
  if (rc != YH_OK) return rc;
  YH_CUDA(cudaMemcpyAsync(h, s->sr_d, sizeof(h), cudaMemcpyDeviceToHost, s->st));
  if (c_phi_h)
    YH_CUDA(cudaMemcpyAsync(c_phi_h + (size_t)6 * done, s->sr_log_d, (size_t)6 * n * sizeof(double),
                            cudaMemcpyDeviceToHost, s->st));
  YH_CUDA(cudaStreamSynchronize(s->st));
  for (int q = 0; q < 3; q++) { s->c[q] = h[YH_SR_C + q]; s->phi[q] = h[YH_SR_PHI + q]; }
  s->have_prev = 0;
  s->raw_input = 0;
  return YH_OK;
}

// contourMode == 1: RD, swap, sAPD every step (main.cu:879-885, 1035), whole batch per launch.
int yh_sim_run_apd(yh_sim *s, int nsteps, const uint8_t *stim_area_h) {
  YH_REQUIRE(s && nsteps >= 0, "bad arguments");
  DevGuard g(s->device);
  const size_t cells = s->n * s->n_sims;
  if (!s->apd[0]) {
    for (int q = 0; q < 6; q++) {
      YH_CUDA(cudaMalloc(&s->apd[q], cells * sizeof(double)));
      YH_CUDA(cudaMemsetAsync(s->apd[q], 0, cells * sizeof(double), s->st));
    }
    YH_CUDA(cudaMalloc(&s->apd_first, cells));
    YH_CUDA(cudaMemsetAsync(s->apd_first, 0, cells, s->st));
    YH_CUDA(cudaMalloc(&s->stim_area, cells));
  }
  if (stim_area_h) {   // the same mask for every sheet; kept for later calls that pass NULL
    for (int z = 0; z < s->n_sims; z++)
      YH_CUDA(cudaMemcpyAsync(s->stim_area + s->n * z, stim_area_h, s->n, cudaMemcpyHostToDevice, s->st));
    s->have_stim_area = 1;
  }
  const bool masked = s->have_stim_area != 0;
  yh_params pb = s->p;   // the whole batch as one tall array for the element-wise APD kernel
  pb.ny = s->p.ny * s->n_sims; pb.ny_global = pb.ny;
  YhK k = yh_make_k(&s->p);
  k.px = s->px; k.py = s->py;
  const bool fused_ok = yh_rd_fast_supported(k, 1) != 0;
  YhApd A{s->apd[0], s->apd[1], s->apd[2], s->apd[3], s->apd[4], s->apd[5], s->apd_first,
          masked ? s->stim_area : nullptr, masked};
  int left = nsteps;
  while (left > 0) {
    if (!s->apd_init || !fused_ok) {
      // full pass: every cell's sAPD / dAPD written, as sAPD_wrapper does every step (main.cu:1035:
      // sAPD_wrapper(..., param.count, gateIn_d.u, gateOut_d.u, ...) AFTER the swap)
      int rc = sim_run_impl(s, 1, 1, nullptr, false);
      if (rc != YH_OK) return rc;
      rc = yh_sapd(&pb, s->count, s->u[s->cur], s->u[s->cur ^ 1], s->apd[0], s->apd[1], s->apd[2], s->apd[3],
                   s->apd[4], s->apd[5], s->apd_first, A.stimArea, A.stimulate, s->st);
      if (rc != YH_OK) return rc;
      s->apd_init = 1;
      left -= 1;
      continue;
    }
    // fused: the state machine runs in the epilogue of the RD kernel, only where u crossed 0.15
    int T = 4;
    while (T > left) T >>= 1;
    const int c = s->cur, o = c ^ 1;
    int rc = yh_launch_rd_fast_paced(k, T, s->u[c], s->v[c], s->u[o], s->v[o], s->n_sims, (long long)s->n,
                                     s->period_d, s->duration_it, s->count, s->raw_input, s->st, &A);
    if (rc != YH_OK) return rc;
    s->cur = o; s->raw_input = 0; s->have_prev = (T == 1); s->count += T; left -= T;
  }
  YH_CUDA(cudaStreamSynchronize(s->st));
  return YH_OK;
}

int yh_sim_get_apd(yh_sim *s, double *apd1_h, double *apd2_h) {
  YH_REQUIRE(s && apd1_h && apd2_h, "null pointer");
  if (!s->apd[0]) { yh_set_error("yh_sim_get_apd: yh_sim_run_apd was never called"); return YH_ERR_UNSUPPORTED; }
  DevGuard g(s->device);
  const size_t bytes = s->n * s->n_sims * sizeof(double);
  YH_CUDA(cudaMemcpy(apd1_h, s->apd[0], bytes, cudaMemcpyDeviceToHost));
  YH_CUDA(cudaMemcpy(apd2_h, s->apd[1], bytes, cudaMemcpyDeviceToHost));
  return YH_OK;
}

int yh_sim_sr_state(yh_sim *s, double c[3], double phi[3], int set) {
  YH_REQUIRE(s && c && phi, "null pointer");
  for (int q = 0; q < 3; q++) {
    if (set) { s->c[q] = c[q]; s->phi[q] = phi[q]; }
    else { c[q] = s->c[q]; phi[q] = s->phi[q]; }
  }
  return YH_OK;
}

int yh_sim_count(const yh_sim *s) { return s ? s->count : -1; }
void *yh_sim_device_u(yh_sim *s) { return s ? s->u[s->cur] : nullptr; }
void *yh_sim_device_v(yh_sim *s) { return s ? s->v[s->cur] : nullptr; }

}  // extern "C"
