// slab.cu -- multi-GPU row-slab driver behind include/yolohtli_slab.h: the reference's host loop
// {reactionDiffusion_wrapper; swapSoA} (main.cu:879-882) for a sheet cut into contiguous row slabs,
// one per GPU.  Host code is C++ like the reference's; the data path is kernels only:
//
//   edge stream (high priority)   wait interior(k-1) | step top band, bottom band | EXCHANGE(k)
//   main stream                   wait bands(k-1)    | step interior rows
//
// EXCHANGE is one kernel: it copies this slab's fresh first / last `halo` owned rows straight into the
// neighbours' ghost rows (16-byte stores through peer mappings = NVLink), the last CTA to finish
// releases the block's sequence number into the neighbours' flag words (st.release.sys) and then
// acquires the two numbers the neighbours release into ours.  The sequence counter lives in device
// memory, so no launch argument changes from block to block and pairs of blocks replay from a CUDA
// graph.  Write-after-read safety: a rank cannot reach exchange k before it has acquired its
// neighbour's exchange k-1, which the neighbour issued after the kernels that read the ghost rows
// exchange k overwrites.
//
// Interior rows are >= halo rows away from the slab edges: they never read a ghost row and overlap
// the exchange.  One HBM pass per block and region (Euler: T <= halo time steps per pass; RK: one
// time step), so no pass of one region reads rows another region is writing.
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <vector>

#include "../../include/yolohtli_slab.h"
#include "yh_common.cuh"

namespace {

constexpr int FLAG_FROM_UP = 0, FLAG_FROM_DOWN = 1, FLAG_STATUS = 2, FLAG_SEQ = 3, FLAG_DONE = 4;
constexpr int FLAG_WORDS = 64;   // 256 bytes
constexpr unsigned HANDLE_MAGIC = 0x59485342u;   // "YHSB"

struct SlabHandle {
  unsigned magic;
  int rank, world, device, own_lo, own_hi, ny_local, nx, pid;
  unsigned long long base_off;           // offset of the slab's block inside the exported allocation
  unsigned long long off[5];             // u0 u1 v0 v1 flags inside the block
  cudaIpcMemHandle_t mem;
};
static_assert(sizeof(SlabHandle) <= YH_SLAB_HANDLE_BYTES, "handle too large");

struct Peer {
  bool present;
  double *u[2], *v[2];
  int *flags;
  int own_lo, own_hi;
  void *ipc_base;                        // cudaIpcOpenMemHandle mapping to close, or NULL
};

struct XchgArgs {
  const double *src_u, *src_v;           // this slab's buffer (stored rows)
  double *dst[4];                        // up.u up.v down.u down.v: first ghost row to write, or NULL
  long long src_off[2];                  // element offset of the rows sent up / down
  long long n;                           // doubles per field and direction (halo * nx)
  int *up_flag, *dn_flag;                // flag words to release into (in the neighbours' memory)
  int *flags;                            // this slab's flag block
};

__device__ __forceinline__ void st_release_sys(int *p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int *p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Peer stores of the fresh edge rows, then release / acquire of the block's sequence number.
__global__ void __launch_bounds__(256) slab_exchange_kernel(const __grid_constant__ XchgArgs a) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthr = (long long)gridDim.x * blockDim.x;
  const bool vec = ((a.n & 1) == 0) && (((a.src_off[0] | a.src_off[1]) & 1) == 0);
  for (int seg = 0; seg < 4; seg++) {
    double *d = a.dst[seg];
    if (!d) continue;
    const double *s = ((seg & 1) ? a.src_v : a.src_u) + a.src_off[seg >> 1];
    if (vec && ((reinterpret_cast<unsigned long long>(d) | reinterpret_cast<unsigned long long>(s)) & 15) == 0) {
      const long long n2 = a.n >> 1;
      for (long long i = tid; i < n2; i += nthr)
        reinterpret_cast<double2 *>(d)[i] = reinterpret_cast<const double2 *>(s)[i];
    } else {
      for (long long i = tid; i < a.n; i += nthr) d[i] = s[i];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x != 0) return;
  unsigned *done = reinterpret_cast<unsigned *>(a.flags + FLAG_DONE);
  if (atomicAdd(done, 1u) != gridDim.x - 1) return;
  // last CTA: every CTA's rows are out (each fenced before it counted itself)
  *done = 0;
  const int seq = a.flags[FLAG_SEQ] + 1;
  __threadfence_system();
  if (a.up_flag) st_release_sys(a.up_flag, seq);
  if (a.dn_flag) st_release_sys(a.dn_flag, seq);
  // a neighbour that never signals must not hang the GPU: ~4 s, then the status word says so
  const long long t0 = clock64();
  for (int w = 0; w < 2; w++) {
    if (!(w == 0 ? a.up_flag : a.dn_flag)) continue;
    const int *f = a.flags + (w == 0 ? FLAG_FROM_UP : FLAG_FROM_DOWN);
    while (ld_acquire_sys(f) < seq) {
      if (clock64() - t0 > 8000000000ll) { a.flags[FLAG_STATUS] = 1; break; }
      __nanosleep(100);
    }
  }
  a.flags[FLAG_SEQ] = seq;
  __threadfence_system();
}

// Sum of the bit patterns of the owned cells mod 2^64 (order independent: one number per sheet, whatever
// the decomposition).
__global__ void slab_checksum_kernel(const unsigned long long *u, const unsigned long long *v, long long n,
                                     unsigned long long *out) {
  unsigned long long su = 0, sv = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    su += u[i]; sv += v[i];
  }
  for (int o = 16; o > 0; o >>= 1) {
    su += __shfl_xor_sync(0xffffffffu, su, o);
    sv += __shfl_xor_sync(0xffffffffu, sv, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(out, su); atomicAdd(out + 1, sv); }
}

struct DevGuard {
  int prev;
  explicit DevGuard(int d) { cudaGetDevice(&prev); cudaSetDevice(d); }
  ~DevGuard() { cudaSetDevice(prev); }
};

struct GraphSlot {
  cudaGraphExec_t exec;
  int n, cur, raw;
};

}  // namespace

struct yh_slab {
  yh_params pg, p;                       // whole sheet / this slab's stored rows
  int rank, world, halo, device, K;
  int j0, j1, g0, g1, ny_local, own_lo, own_hi, nx;
  bool fast;                             // temporally blocked Euler path
  char *block;                           // one allocation: u0 u1 v0 v1 flags
  size_t block_bytes, off[5];
  double *u[2], *v[2];
  int *flags;
  int cur;
  uint8_t *solid, *pat;
  const uint8_t *solid_arg;
  int solid_flags;
  Peer up, down;
  cudaStream_t main, edge, copy;         // copy: host <-> device chunks of the pipelined yh_slab_run_host
  std::vector<cudaEvent_t> ev_chunk;     // per chunk: rows arrived (H2D) / rows final (D2H may start)
  cudaEvent_t ev_int, ev_edge, ev_fork, ev_band;
  bool raw, ghosts_valid, connected;
  long long count;
  unsigned long long *sum_d;
  GraphSlot graphs[4];
  int band;
  // symmetry-reduction mode (yh_slab_group_advance_sr): tangent field, frame velocity, tip list, row sums
  double *sr_vt[2], *sr_adv[2], *sr_rows;
  int *sr_tip_count;
  yh_tip *sr_tip_vec;
  int sr_tip_cap;
  std::vector<cudaEvent_t> tl;           // YH_SLAB_TIMELINE: t0, then 5 events per block (see timeline_dump)
};

struct yh_slab_group {
  std::vector<yh_slab *> m;
  double c[3] = {0.0, 0.0, 0.0}, phi[3] = {0.0, 0.0, 0.0};   // frame velocity and phase (main.cu:902-938)
  long long sr_count = 0;
  std::vector<double> sr_part, sr_sum;
};

namespace {

// YH_SLAB_TIMELINE=<path prefix>: the plain (graph-free) schedule with timing events around every kernel of the
// first 48 blocks of a run, written to <prefix>.<rank> -- the per-rank timeline the profiles/ directory keeps
// (no nsys in this image).  Per block: bands start / bands end / exchange end on the edge stream, interior
// start / interior end on the main stream, microseconds since the start of the run.
const char *timeline_prefix() {
  const char *e = getenv("YH_SLAB_TIMELINE");
  return (e && e[0]) ? e : nullptr;
}
constexpr int TL_BLOCKS = 48;
void tl_mark(yh_slab *s, cudaStream_t st) {
  if (!timeline_prefix() || s->tl.size() >= 1 + 5 * TL_BLOCKS) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  s->tl.push_back(e);
}
void timeline_dump(yh_slab *s) {
  if (!timeline_prefix() || s->tl.size() < 6) return;
  cudaStreamSynchronize(s->edge); cudaStreamSynchronize(s->main);
  char path[512];
  snprintf(path, sizeof(path), "%s.%d", timeline_prefix(), s->rank);
  FILE *f = fopen(path, "w");
  if (f) {
    fprintf(f, "# slab %d of %d, rows [%d, %d), band %d rows; us since the start of the run\n", s->rank, s->world, s->j0, s->j1, s->band);
    fprintf(f, "# block  bands_start  bands_end  exchange_end  interior_start  interior_end\n");
    for (size_t b = 0; 1 + 5 * (b + 1) <= s->tl.size(); b++) {
      fprintf(f, "%3zu", b);
      for (int q = 0; q < 5; q++) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s->tl[0], s->tl[1 + 5 * b + q]);
        fprintf(f, " %10.1f", ms * 1e3);
      }
      fprintf(f, "\n");
    }
    fclose(f);
  }
  for (cudaEvent_t e : s->tl) cudaEventDestroy(e);
  s->tl.clear();
}
// YH_SLAB_WAIT=exchange: the round-1 dependency (interior waits for the whole edge chain of the previous block), A/B
bool wait_for_exchange() {
  static const bool v = [] { const char *e = getenv("YH_SLAB_WAIT"); return e && e[0] == 'e'; }();
  return v;
}

int launch_exchange(yh_slab *s, int buf, cudaStream_t st) {
  XchgArgs a;
  memset(&a, 0, sizeof(a));
  const long long row = s->nx;
  a.src_u = s->u[buf]; a.src_v = s->v[buf];
  a.n = (long long)s->halo * row;
  a.flags = s->flags;
  if (s->up.present) {     // my first halo owned rows are the upper neighbour's lower ghosts
    a.src_off[0] = (long long)s->own_lo * row;
    a.dst[0] = s->up.u[buf] + (long long)s->up.own_hi * row;
    a.dst[1] = s->up.v[buf] + (long long)s->up.own_hi * row;
    a.up_flag = s->up.flags + FLAG_FROM_DOWN;
  }
  if (s->down.present) {   // my last halo owned rows are the lower neighbour's upper ghosts
    a.src_off[1] = (long long)(s->own_hi - s->halo) * row;
    a.dst[2] = s->down.u[buf] + (long long)(s->down.own_lo - s->halo) * row;
    a.dst[3] = s->down.v[buf] + (long long)(s->down.own_lo - s->halo) * row;
    a.dn_flag = s->down.flags + FLAG_FROM_UP;
  }
  const long long pieces = a.n / 2 > 0 ? a.n / 2 : 1;
  int blocks = (int)((pieces + 255) / 256);
  if (blocks > 32) blocks = 32;
  if (blocks < 1) blocks = 1;
  slab_exchange_kernel<<<blocks, 256, 0, st>>>(a);
  YH_LAUNCH_CHECK();
  return YH_OK;
}

int rd_rows_raw(yh_slab *s, int n, int c, int o, int r0, int r1, bool raw, cudaStream_t st);
int rd_rows(yh_slab *s, int n, int c, int o, int r0, int r1, cudaStream_t st) {
  return rd_rows_raw(s, n, c, o, r0, r1, s->raw, st);
}
int rd_rows_raw(yh_slab *s, int n, int c, int o, int r0, int r1, bool raw, cudaStream_t st) {
  if (r1 <= r0) return YH_OK;
  int inB = 0;
  const int flags = (raw ? 0 : YH_RD_INPUT_CANONICAL) | s->solid_flags;
  int rc = yh_rd_advance(&s->p, n, s->fast ? n : 1, flags, s->u[c], s->v[c], s->u[o], s->v[o], s->solid_arg,
                         0, s->nx / 2, s->pg.ny / 2, r0, r1, &inB, st);
  if (rc != YH_OK) return rc;
  if (!inB) { yh_set_error("slab block was not a single pass"); return YH_ERR_UNSUPPORTED; }
  return YH_OK;
}

bool has_neighbours(const yh_slab *s) { return s->up.present || s->down.present; }

// Load the code of every kernel a block of n steps will launch on this slab (the three row ranges may
// take different kernels: tiles for the bands, a streaming kernel for the interior), for raw and for
// library-written input, without launching anything (YH_LAUNCH, yh_common.cuh).
int preload_block(yh_slab *s, int n) {
  const int B = s->band;
  const int top1 = s->up.present ? s->own_lo + B : s->own_lo;
  const int bot0 = s->down.present ? s->own_hi - B : s->own_hi;
  const bool raw0 = s->raw;
  int rc = YH_OK;
  yh_preload_only = 1;
  for (int raw = 0; raw < 2 && rc == YH_OK; raw++) {
    s->raw = raw != 0;
    rc = rd_rows(s, n, 0, 1, s->own_lo, top1, s->main);
    if (rc == YH_OK) rc = rd_rows(s, n, 0, 1, bot0, s->own_hi, s->main);
    if (rc == YH_OK) rc = rd_rows(s, n, 0, 1, top1, bot0, s->main);
  }
  yh_preload_only = 0;
  s->raw = raw0;
  if (rc != YH_OK) return rc;
  cudaFuncAttributes fa;
  YH_CUDA(cudaFuncGetAttributes(&fa, slab_exchange_kernel));
  return YH_OK;
}

// time steps of the next block
int block_steps(const yh_slab *s, int left, int tb) {
  if (!s->fast) return 1;
  int cap = s->halo < left ? s->halo : left;
  const int want = tb ? tb : 4;
  if (want < cap) cap = want;
  int n = 1;
  while (2 * n <= cap && 2 * n <= 4) n *= 2;
  return n;
}

int advance_begin(yh_slab *s) {
  if (s->world > 1 && !s->connected) { yh_set_error("yh_slab_advance before yh_slab_connect"); return YH_ERR_INVALID_ARG; }
  YH_CUDA(cudaEventRecord(s->ev_int, s->main));
  YH_CUDA(cudaEventRecord(s->ev_edge, s->main));
  YH_CUDA(cudaEventRecord(s->ev_band, s->main));
  if (has_neighbours(s) && !s->ghosts_valid) {   // ghost rows of the current state from the neighbours
    YH_CUDA(cudaStreamWaitEvent(s->edge, s->ev_int, 0));
    int rc = launch_exchange(s, s->cur, s->edge);
    if (rc != YH_OK) return rc;
    YH_CUDA(cudaEventRecord(s->ev_edge, s->edge));
    s->ghosts_valid = true;
  }
  return YH_OK;
}

// one block of n time steps: state cur -> cur^1
int advance_block(yh_slab *s, int n) {
  const int c = s->cur, o = c ^ 1;
  int rc;
  if (!has_neighbours(s)) {
    rc = rd_rows(s, n, c, o, s->own_lo, s->own_hi, s->main);
    if (rc != YH_OK) return rc;
  } else {
    const int B = s->band;
    const int top0 = s->own_lo, top1 = s->up.present ? s->own_lo + B : s->own_lo;
    const int bot0 = s->down.present ? s->own_hi - B : s->own_hi, bot1 = s->own_hi;
    YH_CUDA(cudaStreamWaitEvent(s->edge, s->ev_int, 0));     // interior(k-1)
    tl_mark(s, s->edge);
    rc = rd_rows(s, n, c, o, top0, top1, s->edge);
    if (rc != YH_OK) return rc;
    rc = rd_rows(s, n, c, o, bot0, bot1, s->edge);
    if (rc != YH_OK) return rc;
    // the interior of block k reads band rows in state k-1 and overwrites rows the bands of block k-1 read:
    // it depends on those BAND kernels, not on the exchange behind them (whose wait for the neighbours would
    // otherwise sit on the interior's critical path); only the next bands need the exchanged ghost rows, and
    // they follow the exchange in stream order
    YH_CUDA(cudaStreamWaitEvent(s->main, wait_for_exchange() ? s->ev_edge : s->ev_band, 0));    // bands(k-1)
    YH_CUDA(cudaEventRecord(s->ev_band, s->edge));           // bands(k)
    tl_mark(s, s->edge);
    rc = launch_exchange(s, o, s->edge);
    if (rc != YH_OK) return rc;
    tl_mark(s, s->edge);
    tl_mark(s, s->main);
    rc = rd_rows(s, n, c, o, top1, bot0, s->main);
    if (rc != YH_OK) return rc;
    tl_mark(s, s->main);
    YH_CUDA(cudaEventRecord(s->ev_int, s->main));
    YH_CUDA(cudaEventRecord(s->ev_edge, s->edge));
  }
  s->cur = o;
  s->raw = false;
  s->count += n;
  return YH_OK;
}

int advance_end(yh_slab *s) {
  YH_CUDA(cudaStreamWaitEvent(s->main, s->ev_edge, 0));
  return YH_OK;
}

// Two blocks (the ping-pong returns to the same buffers) captured once per {n, parity} and replayed.
int graph_pair(yh_slab *s, int n, cudaGraphExec_t *out) {
  for (auto &g : s->graphs)
    if (g.exec && g.n == n && g.cur == s->cur && g.raw == 0) { *out = g.exec; return YH_OK; }
  GraphSlot *slot = &s->graphs[0];
  for (auto &g : s->graphs)
    if (!g.exec) { slot = &g; break; }
  if (slot->exec) { cudaGraphExecDestroy(slot->exec); slot->exec = nullptr; }
  const int cur0 = s->cur;
  const long long count0 = s->count;
  cudaGraph_t graph = nullptr;
  YH_CUDA(cudaStreamBeginCapture(s->main, cudaStreamCaptureModeThreadLocal));
  int rc = cudaEventRecord(s->ev_fork, s->main) == cudaSuccess ? YH_OK : YH_ERR_CUDA;
  if (rc == YH_OK) rc = cudaStreamWaitEvent(s->edge, s->ev_fork, 0) == cudaSuccess ? YH_OK : YH_ERR_CUDA;
  if (rc == YH_OK) rc = advance_begin(s);
  if (rc == YH_OK) rc = advance_block(s, n);
  if (rc == YH_OK) rc = advance_block(s, n);
  if (rc == YH_OK) rc = advance_end(s);
  cudaError_t ce = cudaStreamEndCapture(s->main, &graph);
  s->cur = cur0; s->count = count0;      // the capture launched nothing
  if (rc != YH_OK || ce != cudaSuccess) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    if (rc == YH_OK) { yh_set_error("graph capture of the slab block pair failed"); rc = YH_ERR_CUDA; }
    return rc;
  }
  cudaGraphExec_t exec = nullptr;
  ce = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) { cudaGetLastError(); yh_set_error("cudaGraphInstantiate failed"); return YH_ERR_CUDA; }
  slot->exec = exec; slot->n = n; slot->cur = cur0; slot->raw = 0;
  *out = exec;
  return YH_OK;
}

bool graphs_wanted() {
  const char *f = getenv("YH_SLAB_GRAPHS");
  return !(f && f[0] == '0') && !timeline_prefix();
}

// The whole run of one slab or, in lock step, of all slabs of a group (breadth first per block, so no
// slab's queue runs ahead of a neighbour whose work is not enqueued yet).
int advance_all(yh_slab *const *m, int count, int nsteps, int tb) {
  int rc;
  // every kernel of the run resident before the first exchange kernel can be waiting for a neighbour
  for (int left = nsteps; left > 0;) {
    const int n = block_steps(m[0], left, tb);
    for (int q = 0; q < count; q++) {
      DevGuard g(m[q]->device);
      rc = preload_block(m[q], n);
      if (rc != YH_OK) return rc;
    }
    left -= (left / n) * n;
  }
  for (int q = 0; q < count; q++) {
    DevGuard g(m[q]->device);
    rc = advance_begin(m[q]);
    if (rc != YH_OK) return rc;
    if (has_neighbours(m[q]) && m[q]->tl.empty()) tl_mark(m[q], m[q]->main);
  }
  int left = nsteps;
  while (left > 0) {
    const int n = block_steps(m[0], left, tb);
    const int same = left / n;             // consecutive blocks of this length
    int plain = same;
    // steady state from graphs: two plain blocks first (raw input, first use of the kernel variants)
    if (graphs_wanted() && same >= 8 && has_neighbours(m[0])) plain = 2 + (same - 2) % 2;
    for (int b = 0; b < plain; b++) {
      for (int q = 0; q < count; q++) {
        DevGuard g(m[q]->device);
        rc = advance_block(m[q], n);
        if (rc != YH_OK) return rc;
      }
      if (getenv("YH_SLAB_DEBUG_SYNC"))    // debugging aid: host and devices in lock step
        for (int q = 0; q < count; q++) { cudaStreamSynchronize(m[q]->edge); cudaStreamSynchronize(m[q]->main); }
    }
    const int pairs = (same - plain) / 2;
    if (pairs > 0) {
      std::vector<cudaGraphExec_t> ex(count, nullptr);
      bool ok = true;
      for (int q = 0; q < count; q++) {
        DevGuard g(m[q]->device);
        rc = advance_end(m[q]);            // join the edge stream: the graph runs on main
        if (rc != YH_OK) return rc;
        if (graph_pair(m[q], n, &ex[q]) != YH_OK) ok = false;
      }
      if (ok) {
        for (int b = 0; b < pairs; b++)
          for (int q = 0; q < count; q++) {
            DevGuard g(m[q]->device);
            YH_CUDA(cudaGraphLaunch(ex[q], m[q]->main));
          }
        for (int q = 0; q < count; q++) {
          DevGuard g(m[q]->device);
          m[q]->count += (long long)2 * n * pairs;
          rc = advance_begin(m[q]);        // events of the plain schedule again
          if (rc != YH_OK) return rc;
        }
      } else {                             // capture refused: plain launches
        for (int b = 0; b < 2 * pairs; b++)
          for (int q = 0; q < count; q++) {
            DevGuard g(m[q]->device);
            rc = advance_block(m[q], n);
            if (rc != YH_OK) return rc;
          }
      }
    }
    left -= same * n;
  }
  for (int q = 0; q < count; q++) {
    DevGuard g(m[q]->device);
    rc = advance_end(m[q]);
    if (rc != YH_OK) return rc;
  }
  if (timeline_prefix())
    for (int q = 0; q < count; q++) {
      DevGuard g(m[q]->device);
      if (m[q]->tl.size() >= 1 + 5 * TL_BLOCKS) timeline_dump(m[q]);
    }
  return YH_OK;
}


// ---- yh_slab_run_host with the copies hidden behind the time steps ----------------------------------------
// Host -> device, nsteps, device -> host is what a caller of the reference does (main.cu:470 ... 519).  The
// copies of a 16384^2 sheet take 2 x 75 ms over PCIe against 730 ms of time steps on one GPU, and 2 x 35 ms
// against 100 ms on each of 8 GPUs that share the host's memory -- so the owned rows travel in C chunks and
// the time steps start on a chunk as soon as it has arrived (and end on a chunk while the others still run):
//
//   level b = state after b blocks of n time steps; a block needs h = n * timeIntOrder rows of the previous
//   level on either side.  Chunk c at level b covers rows [X_c - h*b, X_{c+1} - h*b): every level the
//   boundaries move UP by h rows, so block (c, b) reads only (c, b-1) and (c-1, b-1) -- chunks that arrived
//   EARLIER.  Prologue: after chunk c has arrived it runs P blocks (while chunk c+1 is in flight).  The rows
//   of a slab edge that faces a neighbour cannot follow without the neighbour's rows: that edge recedes by h
//   rows per level and the wedge left behind (h*P rows) is caught up afterwards, level by level with one
//   halo exchange each, like ordinary blocks on a few rows.  Then all rows are at level P and the run
//   continues with whole-slab blocks (advance_all).  Epilogue: the mirror image -- chunk after chunk runs
//   its last E blocks and its rows start their way back to the host while the next chunk computes; the edge
//   wedges are finished last.
// Every cell update is the same function of the same inputs as in the plain schedule: results are bit-identical
// (tests/slab_driver.cu `pipe` mode, tests/test_gpu_slab_driver.py).  YH_SLAB_PIPE=0 switches it off.
struct PipePlan {
  int n, h, B, P, E, C, S;
};

int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}

// What the plan depends on -- host arithmetic only, so that it can be checked without a device
// (yh_slab_pipeline_plan / _region, tests/test_slab_pipeline_plan.py).  Rows are LOCAL rows of the slab's arrays.
struct PipeGeom {
  bool fast, solid, has_up, has_down;
  int halo, K, own_lo, own_hi, ny, world;
};

PipeGeom geom_of(const yh_slab *s) {
  PipeGeom g;
  g.fast = s->fast; g.solid = s->solid_arg != nullptr; g.has_up = s->rank > 0; g.has_down = s->rank < s->world - 1;
  g.halo = s->halo; g.K = s->K; g.own_lo = s->own_lo; g.own_hi = s->own_hi; g.ny = s->pg.ny; g.world = s->world;
  return g;
}

int block_steps_g(bool fast, int halo, int left, int tb) {   // block_steps without a slab
  if (!fast) return 1;
  int cap = halo < left ? halo : left;
  const int want = tb ? tb : 4;
  if (want < cap) cap = want;
  int n = 1;
  while (2 * n <= cap && 2 * n <= 4) n *= 2;
  return n;
}

bool plan_geom(const PipeGeom &g, int nsteps, int tb, PipePlan *pl) {
  if (env_int("YH_SLAB_PIPE", 1) == 0) return false;
  if (g.solid) return false;                             // masks: plain schedule (untested here)
  const int n = block_steps_g(g.fast, g.halo, nsteps, tb);
  const int h = n * g.K;
  if (h > g.halo || nsteps < n) return false;
  const int own = g.own_hi - g.own_lo;
  // every rank must reach the same plan on its own (one exchange per wedge level on both sides of a slab
  // boundary): chunk count and level count come from the SMALLEST slab height of the partition, not from this slab's
  const int own_min = g.ny / g.world;
  int C = env_int("YH_SLAB_PIPE_CHUNKS", own_min >= 8192 ? 8 : 4);
  if (C < 2 || own_min < 64 * C) return false;
  const int S = (own + C - 1) / C;
  if (own - (C - 1) * S < 1) return false;
  // blocks a chunk runs while the next chunk is in flight: copy time of a chunk / time of one block on it.
  // Euler (4 steps per block at ~370 Gcell/s) against ~50 GB/s of pinned copies: ~28; RK4: ~10
  int P = env_int("YH_SLAB_PIPE_LEVELS", g.fast ? 28 : 10);
  const int cap = ((own_min + C - 1) / C - 16) / (2 * h);   // chunk 0 keeps >= 16 rows: S - h*P (shift) - h*P (wedge)
  if (P > cap) P = cap;
  const int B = nsteps / n;
  if (P > (B - 1) / 2) P = (B - 1) / 2;
  if (P < 2) return false;
  pl->n = n; pl->h = h; pl->B = B; pl->P = P; pl->E = P; pl->C = C; pl->S = S;
  return true;
}

bool plan_pipeline(const yh_slab *s, int nsteps, int tb, PipePlan *pl) { return plan_geom(geom_of(s), nsteps, tb, pl); }

// rows of chunk c at (relative) level b >= 1
void region_geom(const PipeGeom &g, const PipePlan &pl, int c, int b, int *r0, int *r1) {
  const int X0 = g.own_lo + c * pl.S;
  const int X1 = (c == pl.C - 1) ? g.own_hi : g.own_lo + (c + 1) * pl.S;
  *r0 = (c == 0) ? (g.has_up ? g.own_lo + pl.h * b : g.own_lo) : X0 - pl.h * b;
  *r1 = (c == pl.C - 1) ? (g.has_down ? g.own_hi - pl.h * b : g.own_hi) : X1 - pl.h * b;
}

void pipe_region(const yh_slab *s, const PipePlan &pl, int c, int b, int *r0, int *r1) {
  region_geom(geom_of(s), pl, c, b, r0, r1);
}

// levels 1 .. L of every chunk, chunk-major; base = buffer that holds level 0.  wait_chunks: the prologue (chunk
// c waits for its rows); done_chunks: the epilogue (rows of chunk c at level L go home once they are final).
int pipe_chunks(yh_slab *s, const PipePlan &pl, int L, int base, bool raw0, bool prologue, double *u_out_h, double *v_out_h) {
  const size_t row = (size_t)s->nx;
  for (int c = 0; c < pl.C; c++) {
    if (prologue) YH_CUDA(cudaStreamWaitEvent(s->main, s->ev_chunk[c], 0));
    for (int b = 1; b <= L; b++) {
      int r0, r1;
      pipe_region(s, pl, c, b, &r0, &r1);
      int rc = rd_rows_raw(s, pl.n, (base + b - 1) & 1, (base + b) & 1, r0, r1, raw0 && b == 1, s->main);
      if (rc != YH_OK) return rc;
    }
    if (!prologue) {
      int r0, r1;
      pipe_region(s, pl, c, L, &r0, &r1);
      YH_CUDA(cudaEventRecord(s->ev_chunk[c], s->main));
      YH_CUDA(cudaStreamWaitEvent(s->copy, s->ev_chunk[c], 0));
      const int fb = (base + L) & 1;
      const size_t ho = (size_t)(r0 - s->own_lo) * row, bytes = (size_t)(r1 - r0) * row * sizeof(double);
      YH_CUDA(cudaMemcpyAsync(u_out_h + ho, s->u[fb] + (size_t)r0 * row, bytes, cudaMemcpyDeviceToHost, s->copy));
      YH_CUDA(cudaMemcpyAsync(v_out_h + ho, s->v[fb] + (size_t)r0 * row, bytes, cudaMemcpyDeviceToHost, s->copy));
    }
  }
  return YH_OK;
}

// wedge catch-up, level j of L (all slabs of the call in lock step): exchange of the level j-1 edge rows, then the
// top / bottom wedge rows [lo, lo + h*j) / [hi - h*j, hi) go from level j-1 to j.
int pipe_wedge_level(yh_slab *s, const PipePlan &pl, int j, int base, bool raw0) {
  if (!has_neighbours(s)) return YH_OK;
  int rc = launch_exchange(s, (base + j - 1) & 1, s->main);
  if (rc != YH_OK) return rc;
  const bool raw = raw0 && j == 1;
  if (s->up.present) {
    rc = rd_rows_raw(s, pl.n, (base + j - 1) & 1, (base + j) & 1, s->own_lo, s->own_lo + pl.h * j, raw, s->main);
    if (rc != YH_OK) return rc;
  }
  if (s->down.present) {
    rc = rd_rows_raw(s, pl.n, (base + j - 1) & 1, (base + j) & 1, s->own_hi - pl.h * j, s->own_hi, raw, s->main);
    if (rc != YH_OK) return rc;
  }
  return YH_OK;
}

int pipe_preload(yh_slab *s, const PipePlan &pl) {
  // code of every kernel variant the chunk and wedge passes can take, loaded before any exchange kernel waits
  int rc = YH_OK;
  yh_preload_only = 1;
  for (int raw = 0; raw < 2 && rc == YH_OK; raw++) {
    int r0, r1;
    for (int c = 0; c < pl.C && rc == YH_OK; c++) {
      pipe_region(s, pl, c, 1, &r0, &r1);
      rc = rd_rows_raw(s, pl.n, 0, 1, r0, r1, raw != 0, s->main);
      pipe_region(s, pl, c, pl.P, &r0, &r1);
      if (rc == YH_OK) rc = rd_rows_raw(s, pl.n, 0, 1, r0, r1, raw != 0, s->main);
    }
    for (int j = 1; j <= pl.P && rc == YH_OK; j++)
      rc = rd_rows_raw(s, pl.n, 0, 1, s->own_lo, s->own_lo + pl.h * j, raw != 0, s->main);
  }
  yh_preload_only = 0;
  if (rc != YH_OK) return rc;
  cudaFuncAttributes fa;
  YH_CUDA(cudaFuncGetAttributes(&fa, slab_exchange_kernel));
  return YH_OK;
}

// u_in / u_out ...: per slab, the host arrays of its OWNED rows
int run_host_pipelined(yh_slab *const *m, int count, const double *const *u_in, const double *const *v_in,
                       double *const *u_out, double *const *v_out, int nsteps, int tb, bool *done) {
  *done = false;
  PipePlan pl;
  if (!plan_pipeline(m[0], nsteps, tb, &pl)) return YH_OK;
  for (int q = 1; q < count; q++) {       // one plan for all slabs of the call (lock step): the most restrictive
    PipePlan o;
    if (!plan_pipeline(m[q], nsteps, tb, &o) || o.n != pl.n || o.C != pl.C) return YH_OK;
    if (o.P < pl.P) { pl.P = o.P; pl.E = o.P; }
  }
  int rc;
  std::vector<int> base(count);
  for (int q = 0; q < count; q++) {
    yh_slab *s = m[q];
    DevGuard g(s->device);
    if (s->world > 1 && !s->connected) { yh_set_error("yh_slab_run_host before yh_slab_connect"); return YH_ERR_INVALID_ARG; }
    while ((int)s->ev_chunk.size() < pl.C) {
      cudaEvent_t e;
      YH_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      s->ev_chunk.push_back(e);
    }
    rc = pipe_preload(s, pl);
    if (rc != YH_OK) return rc;
    // ---- host -> device, chunk by chunk (the copy stream follows everything enqueued on main so far) ----
    YH_CUDA(cudaEventRecord(s->ev_fork, s->main));
    YH_CUDA(cudaStreamWaitEvent(s->copy, s->ev_fork, 0));
    const size_t row = (size_t)s->nx;
    for (int c = 0; c < pl.C; c++) {
      const int X0 = s->own_lo + c * pl.S, X1 = (c == pl.C - 1) ? s->own_hi : X0 + pl.S;
      const size_t ho = (size_t)(X0 - s->own_lo) * row, bytes = (size_t)(X1 - X0) * row * sizeof(double);
      YH_CUDA(cudaMemcpyAsync(s->u[s->cur] + (size_t)X0 * row, u_in[q] + ho, bytes, cudaMemcpyHostToDevice, s->copy));
      YH_CUDA(cudaMemcpyAsync(s->v[s->cur] + (size_t)X0 * row, v_in[q] + ho, bytes, cudaMemcpyHostToDevice, s->copy));
      YH_CUDA(cudaEventRecord(s->ev_chunk[c], s->copy));
    }
    base[q] = s->cur;
  }
  // ---- prologue: P levels per chunk, then the wedges ----
  for (int q = 0; q < count; q++) {
    DevGuard g(m[q]->device);
    rc = pipe_chunks(m[q], pl, pl.P, base[q], true, true, nullptr, nullptr);
    if (rc != YH_OK) return rc;
  }
  for (int j = 1; j <= pl.P; j++)
    for (int q = 0; q < count; q++) {
      DevGuard g(m[q]->device);
      rc = pipe_wedge_level(m[q], pl, j, base[q], true);
      if (rc != YH_OK) return rc;
    }
  for (int q = 0; q < count; q++) {
    yh_slab *s = m[q];
    s->cur = (base[q] + pl.P) & 1; s->raw = false; s->ghosts_valid = false; s->count += (long long)pl.P * pl.n;
  }
  // ---- whole-slab blocks ----
  const int mid = nsteps - (pl.P + pl.E) * pl.n;
  if (mid > 0) {
    rc = advance_all(m, count, mid, tb);
    if (rc != YH_OK) return rc;
  }
  // ---- epilogue: E levels per chunk with its rows leaving as soon as they are final, then the wedges ----
  for (int q = 0; q < count; q++) {
    DevGuard g(m[q]->device);
    base[q] = m[q]->cur;
    rc = pipe_chunks(m[q], pl, pl.E, base[q], false, false, u_out[q], v_out[q]);
    if (rc != YH_OK) return rc;
  }
  for (int j = 1; j <= pl.E; j++)
    for (int q = 0; q < count; q++) {
      DevGuard g(m[q]->device);
      rc = pipe_wedge_level(m[q], pl, j, base[q], false);
      if (rc != YH_OK) return rc;
    }
  for (int q = 0; q < count; q++) {
    yh_slab *s = m[q];
    DevGuard g(s->device);
    const int fb = (base[q] + pl.E) & 1;
    const size_t row = (size_t)s->nx;
    YH_CUDA(cudaEventRecord(s->ev_fork, s->main));
    YH_CUDA(cudaStreamWaitEvent(s->copy, s->ev_fork, 0));
    for (int w = 0; w < 2; w++) {
      if (!(w == 0 ? s->up.present : s->down.present)) continue;
      const int r0 = w == 0 ? s->own_lo : s->own_hi - pl.h * pl.E, r1 = r0 + pl.h * pl.E;
      const size_t ho = (size_t)(r0 - s->own_lo) * row, bytes = (size_t)(r1 - r0) * row * sizeof(double);
      YH_CUDA(cudaMemcpyAsync(u_out[q] + ho, s->u[fb] + (size_t)r0 * row, bytes, cudaMemcpyDeviceToHost, s->copy));
      YH_CUDA(cudaMemcpyAsync(v_out[q] + ho, s->v[fb] + (size_t)r0 * row, bytes, cudaMemcpyDeviceToHost, s->copy));
    }
    s->cur = fb; s->ghosts_valid = false; s->count += (long long)pl.E * pl.n;
    // main joins the copy stream: later work on this slab follows the copies
    YH_CUDA(cudaEventRecord(s->ev_fork, s->copy));
    YH_CUDA(cudaStreamWaitEvent(s->main, s->ev_fork, 0));
  }
  *done = true;
  return YH_OK;
}

}  // namespace

extern "C" {

int yh_slab_partition(int ny_global, int world, int rank, int *j0, int *j1) {
  if (!j0 || !j1 || world < 1 || rank < 0 || rank >= world || ny_global < world) {
    yh_set_error("yh_slab_partition: bad arguments");
    return YH_ERR_INVALID_ARG;
  }
  const int base = ny_global / world, rem = ny_global % world;
  *j0 = rank * base + (rank < rem ? rank : rem);
  *j1 = *j0 + base + (rank < rem ? 1 : 0);
  return YH_OK;
}

int yh_slab_create(yh_slab **out, const yh_params *pg, int rank, int world, int halo, int device) {
  int rc = yh_check_device();
  if (rc != YH_OK) return rc;
  YH_REQUIRE(out && pg, "null pointer");
  YH_REQUIRE(pg->jg0 == 0 && pg->ny_global == pg->ny, "p_global must describe the whole sheet");
  YH_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank / world");
  YH_REQUIRE(device >= 0 && device < yh_device_count(), "no such device");
  YH_REQUIRE(pg->timeIntOrder == 1 || pg->timeIntOrder == 2 || pg->timeIntOrder == 4, "timeIntOrder must be 1, 2 or 4");
  YH_REQUIRE(halo >= pg->timeIntOrder && halo >= 1, "halo must be >= timeIntOrder");
  if (world > 1 && pg->anisotropy && pg->neumannBC && !pg->solidSwitch) {
    // the anisotropic no-flux corner terms read rows j+-2 at the x edges (reactionDiffusion.cu:290-304)
    yh_set_error("anisotropy with no-flux boundaries is not supported on row slabs");
    return YH_ERR_UNSUPPORTED;
  }
  DevGuard g(device);
  yh_slab *s = new yh_slab();
  memset(&s->up, 0, sizeof(Peer)); memset(&s->down, 0, sizeof(Peer));
  memset(s->graphs, 0, sizeof(s->graphs));
  s->pg = *pg; s->rank = rank; s->world = world; s->halo = halo; s->device = device;
  s->K = pg->timeIntOrder; s->nx = pg->nx;
  yh_slab_partition(pg->ny, world, rank, &s->j0, &s->j1);
  if (world > 1 && s->j1 - s->j0 < 2 * halo) {
    delete s;
    yh_set_error("slab thinner than two halos");
    return YH_ERR_INVALID_ARG;
  }
  s->g0 = s->j0 - halo > 0 ? s->j0 - halo : 0;
  s->g1 = s->j1 + halo < pg->ny ? s->j1 + halo : pg->ny;
  s->ny_local = s->g1 - s->g0;
  s->own_lo = s->j0 - s->g0; s->own_hi = s->j1 - s->g0;
  s->p = *pg; s->p.ny = s->ny_local; s->p.ny_global = pg->ny; s->p.jg0 = s->g0;
  const int own = s->j1 - s->j0;
  int B = own / 4 < 128 ? own / 4 : 128;
  if (const char *f = getenv("YH_SLAB_BAND")) B = atoi(f);   // tuning hook
  if (B > own / 2) B = own / 2;
  if (B < halo) B = halo;
  s->band = B;
  YhK k = yh_make_k(&s->p);
  s->fast = pg->solidSwitch ? yh_rd_fast_solid_supported(k, 1) != 0 : yh_rd_fast_supported(k, 1) != 0;
  const size_t arr = (((size_t)s->ny_local * s->nx * sizeof(double)) + 255) & ~(size_t)255;
  for (int q = 0; q < 4; q++) s->off[q] = arr * q;
  s->off[4] = arr * 4;
  s->block_bytes = arr * 4 + FLAG_WORDS * sizeof(int);
  if (cudaMalloc(&s->block, s->block_bytes) != cudaSuccess) {
    yh_set_error("yh_slab_create: cudaMalloc of %zu bytes failed", s->block_bytes);
    cudaGetLastError();
    delete s;
    return YH_ERR_CUDA;
  }
  cudaMemset(s->block, 0, s->block_bytes);
  s->u[0] = reinterpret_cast<double *>(s->block + s->off[0]);
  s->u[1] = reinterpret_cast<double *>(s->block + s->off[1]);
  s->v[0] = reinterpret_cast<double *>(s->block + s->off[2]);
  s->v[1] = reinterpret_cast<double *>(s->block + s->off[3]);
  s->flags = reinterpret_cast<int *>(s->block + s->off[4]);
  s->cur = 0; s->solid = nullptr; s->pat = nullptr; s->solid_arg = nullptr; s->solid_flags = 0;
  s->raw = true; s->ghosts_valid = false; s->connected = (world == 1); s->count = 0;
  s->sr_vt[0] = s->sr_vt[1] = s->sr_adv[0] = s->sr_adv[1] = s->sr_rows = nullptr;
  s->sr_tip_count = nullptr; s->sr_tip_vec = nullptr; s->sr_tip_cap = 0;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);   // hi = greatest priority (numerically lowest)
  cudaStreamCreateWithPriority(&s->main, cudaStreamNonBlocking, lo);
  cudaStreamCreateWithPriority(&s->edge, cudaStreamNonBlocking, hi);
  cudaStreamCreateWithPriority(&s->copy, cudaStreamNonBlocking, lo);
  cudaEventCreateWithFlags(&s->ev_int, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&s->ev_edge, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&s->ev_band, cudaEventDisableTiming);
  cudaMalloc(&s->sum_d, 2 * sizeof(unsigned long long));
  cudaDeviceSynchronize();
  if (cudaGetLastError() != cudaSuccess) { yh_set_error("yh_slab_create: stream / event setup failed"); return YH_ERR_CUDA; }
  *out = s;
  return YH_OK;
}

int yh_slab_destroy(yh_slab *s) {
  if (!s) return YH_OK;
  DevGuard g(s->device);
  cudaStreamSynchronize(s->main);
  cudaStreamSynchronize(s->edge);
  for (auto &gr : s->graphs)
    if (gr.exec) cudaGraphExecDestroy(gr.exec);
  if (s->up.ipc_base) cudaIpcCloseMemHandle(s->up.ipc_base);
  if (s->down.ipc_base) cudaIpcCloseMemHandle(s->down.ipc_base);
  cudaFree(s->block); cudaFree(s->solid); cudaFree(s->pat); cudaFree(s->sum_d);
  cudaFree(s->sr_vt[0]); cudaFree(s->sr_vt[1]); cudaFree(s->sr_adv[0]); cudaFree(s->sr_adv[1]); cudaFree(s->sr_rows);
  cudaFree(s->sr_tip_count); cudaFree(s->sr_tip_vec);
  cudaStreamDestroy(s->main); cudaStreamDestroy(s->edge); cudaStreamDestroy(s->copy);
  for (cudaEvent_t e : s->ev_chunk) cudaEventDestroy(e);
  cudaEventDestroy(s->ev_int); cudaEventDestroy(s->ev_edge); cudaEventDestroy(s->ev_fork); cudaEventDestroy(s->ev_band);
  cudaGetLastError();
  delete s;
  return YH_OK;
}

int yh_slab_layout(const yh_slab *s, int *j0, int *j1, int *g0, int *g1) {
  YH_REQUIRE(s != nullptr, "null slab");
  if (j0) *j0 = s->j0;
  if (j1) *j1 = s->j1;
  if (g0) *g0 = s->g0;
  if (g1) *g1 = s->g1;
  return YH_OK;
}

int yh_slab_export(yh_slab *s, void *handle) {
  YH_REQUIRE(s && handle, "null pointer");
  DevGuard g(s->device);
  SlabHandle h;
  memset(&h, 0, sizeof(h));
  h.magic = HANDLE_MAGIC; h.rank = s->rank; h.world = s->world; h.device = s->device;
  h.own_lo = s->own_lo; h.own_hi = s->own_hi; h.ny_local = s->ny_local; h.nx = s->nx; h.pid = (int)getpid();
  for (int q = 0; q < 5; q++) h.off[q] = s->off[q];
  // the IPC handle names the ALLOCATION the block lives in; cudaMalloc may place a block inside a larger
  // reservation, so the offset from the allocation base travels with it (driver entry point resolved at
  // run time: the library has no link-time dependency on libcuda)
  typedef int (*GetRange)(unsigned long long *, size_t *, unsigned long long);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  h.base_off = 0;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn) {
    unsigned long long base = 0;
    size_t size = 0;
    if (reinterpret_cast<GetRange>(fn)(&base, &size, (unsigned long long)(uintptr_t)s->block) == 0)
      h.base_off = (unsigned long long)(uintptr_t)s->block - base;
  } else {
    cudaGetLastError();
  }
  YH_CUDA(cudaIpcGetMemHandle(&h.mem, s->block));
  memset(handle, 0, YH_SLAB_HANDLE_BYTES);
  memcpy(handle, &h, sizeof(h));
  return YH_OK;
}

static int open_peer(yh_slab *s, const void *handle, Peer *p, int expect_rank) {
  SlabHandle h;
  memcpy(&h, handle, sizeof(h));
  YH_REQUIRE(h.magic == HANDLE_MAGIC, "not a slab handle");
  YH_REQUIRE(h.rank == expect_rank && h.world == s->world && h.nx == s->nx, "handle of the wrong rank / sheet");
  void *base = nullptr;
  YH_CUDA(cudaIpcOpenMemHandle(&base, h.mem, cudaIpcMemLazyEnablePeerAccess));
  char *blk = static_cast<char *>(base) + h.base_off;
  p->present = true; p->ipc_base = base;
  p->u[0] = reinterpret_cast<double *>(blk + h.off[0]); p->u[1] = reinterpret_cast<double *>(blk + h.off[1]);
  p->v[0] = reinterpret_cast<double *>(blk + h.off[2]); p->v[1] = reinterpret_cast<double *>(blk + h.off[3]);
  p->flags = reinterpret_cast<int *>(blk + h.off[4]);
  p->own_lo = h.own_lo; p->own_hi = h.own_hi;
  return YH_OK;
}

int yh_slab_connect(yh_slab *s, const void *handle_up, const void *handle_down) {
  YH_REQUIRE(s != nullptr, "null slab");
  YH_REQUIRE((handle_up != nullptr) == (s->rank > 0), "handle_up must be given exactly when rank > 0");
  YH_REQUIRE((handle_down != nullptr) == (s->rank < s->world - 1), "handle_down must be given exactly when rank < world-1");
  DevGuard g(s->device);
  int rc;
  if (handle_up && (rc = open_peer(s, handle_up, &s->up, s->rank - 1)) != YH_OK) return rc;
  if (handle_down && (rc = open_peer(s, handle_down, &s->down, s->rank + 1)) != YH_OK) return rc;
  s->connected = true;
  return YH_OK;
}

static int local_peer(yh_slab *s, yh_slab *o, Peer *p) {
  if (o->device != s->device) {
    int rc = yh_enable_peer_access(o->device);
    if (rc != YH_OK) return rc;
  }
  p->present = true; p->ipc_base = nullptr;
  p->u[0] = o->u[0]; p->u[1] = o->u[1]; p->v[0] = o->v[0]; p->v[1] = o->v[1];
  p->flags = o->flags; p->own_lo = o->own_lo; p->own_hi = o->own_hi;
  return YH_OK;
}

int yh_slab_connect_local(yh_slab *s, yh_slab *up, yh_slab *down) {
  YH_REQUIRE(s != nullptr, "null slab");
  YH_REQUIRE((up != nullptr) == (s->rank > 0) && (down != nullptr) == (s->rank < s->world - 1),
             "neighbours must be given exactly where they exist");
  YH_REQUIRE((!up || up->rank == s->rank - 1) && (!down || down->rank == s->rank + 1), "wrong neighbour rank");
  DevGuard g(s->device);
  int rc;
  if (up && (rc = local_peer(s, up, &s->up)) != YH_OK) return rc;
  if (down && (rc = local_peer(s, down, &s->down)) != YH_OK) return rc;
  s->connected = true;
  return YH_OK;
}

int yh_slab_set_state(yh_slab *s, const double *u_h, const double *v_h, int with_ghosts) {
  YH_REQUIRE(s && u_h && v_h, "null pointer");
  DevGuard g(s->device);
  const size_t row = (size_t)s->nx;
  const size_t first = with_ghosts ? 0 : (size_t)s->own_lo;
  const size_t rows = with_ghosts ? (size_t)s->ny_local : (size_t)(s->own_hi - s->own_lo);
  YH_CUDA(cudaMemcpyAsync(s->u[s->cur] + first * row, u_h, rows * row * sizeof(double), cudaMemcpyHostToDevice, s->main));
  YH_CUDA(cudaMemcpyAsync(s->v[s->cur] + first * row, v_h, rows * row * sizeof(double), cudaMemcpyHostToDevice, s->main));
  s->raw = true;
  s->ghosts_valid = with_ghosts != 0 || s->world == 1;
  return YH_OK;
}

int yh_slab_get_state(yh_slab *s, double *u_h, double *v_h) {
  YH_REQUIRE(s && u_h && v_h, "null pointer");
  DevGuard g(s->device);
  const size_t row = (size_t)s->nx, rows = (size_t)(s->own_hi - s->own_lo);
  YH_CUDA(cudaMemcpyAsync(u_h, s->u[s->cur] + (size_t)s->own_lo * row, rows * row * sizeof(double), cudaMemcpyDeviceToHost, s->main));
  YH_CUDA(cudaMemcpyAsync(v_h, s->v[s->cur] + (size_t)s->own_lo * row, rows * row * sizeof(double), cudaMemcpyDeviceToHost, s->main));
  return yh_slab_sync(s);
}

int yh_slab_set_solid(yh_slab *s, const uint8_t *mask_h) {
  YH_REQUIRE(s && mask_h, "null pointer");
  DevGuard g(s->device);
  const size_t n = (size_t)s->ny_local * s->nx;
  if (!s->solid) YH_CUDA(cudaMalloc(&s->solid, n));
  YH_CUDA(cudaMemcpyAsync(s->solid, mask_h, n, cudaMemcpyHostToDevice, s->main));
  s->solid_arg = s->solid; s->solid_flags = 0;
  YhK k = yh_make_k(&s->p);
  if (yh_rd_fast_solid_supported(k, 1)) {   // masked Euler: neighbourhood patterns once (the mask is fixed)
    if (!s->pat) YH_CUDA(cudaMalloc(&s->pat, n));
    int rc = yh_rd_solid_patterns(k, s->solid, s->pat, s->main);
    if (rc != YH_OK) return rc;
    s->solid_arg = s->pat; s->solid_flags = YH_RD_SOLID_IS_PATTERNS;
  }
  YH_CUDA(cudaStreamSynchronize(s->main));
  return YH_OK;
}

void *yh_slab_device_u(yh_slab *s) { return s ? s->u[s->cur] : nullptr; }
void *yh_slab_device_v(yh_slab *s) { return s ? s->v[s->cur] : nullptr; }
void *yh_slab_stream(yh_slab *s) { return s ? (void *)s->main : nullptr; }

int yh_slab_advance(yh_slab *s, int nsteps, int tb_steps) {
  YH_REQUIRE(s && nsteps >= 0, "bad arguments");
  YH_REQUIRE(tb_steps == 0 || tb_steps == 1 || tb_steps == 2 || tb_steps == 4, "tb_steps must be 0, 1, 2 or 4");
  YH_REQUIRE(!s->pg.solidSwitch || s->solid_arg, "solidSwitch set: call yh_slab_set_solid first");
  if (nsteps == 0) return YH_OK;
  return advance_all(&s, 1, nsteps, tb_steps);
}

int yh_slab_sync(yh_slab *s) {
  YH_REQUIRE(s != nullptr, "null slab");
  DevGuard g(s->device);
  YH_CUDA(cudaStreamSynchronize(s->edge));
  YH_CUDA(cudaStreamSynchronize(s->main));
  int f[8] = {0};
  YH_CUDA(cudaMemcpy(f, s->flags, sizeof(f), cudaMemcpyDeviceToHost));
  if (f[FLAG_STATUS] != 0) {
    yh_set_error("slab %d: a neighbour never released its halo rows (flag wait timed out; from_up %d from_down %d "
                 "own sequence %d arrivals %d); ghost rows are stale", s->rank, f[FLAG_FROM_UP], f[FLAG_FROM_DOWN],
                 f[FLAG_SEQ], f[FLAG_DONE]);
    return YH_ERR_CUDA;
  }
  return YH_OK;
}

int yh_slab_checksum(yh_slab *s, unsigned long long *sum_u, unsigned long long *sum_v) {
  YH_REQUIRE(s && sum_u && sum_v, "null pointer");
  DevGuard g(s->device);
  YH_CUDA(cudaMemsetAsync(s->sum_d, 0, 2 * sizeof(unsigned long long), s->main));
  const long long n = (long long)(s->own_hi - s->own_lo) * s->nx;
  const size_t o = (size_t)s->own_lo * s->nx;
  slab_checksum_kernel<<<148 * 4, 256, 0, s->main>>>(reinterpret_cast<const unsigned long long *>(s->u[s->cur] + o),
                                                     reinterpret_cast<const unsigned long long *>(s->v[s->cur] + o), n, s->sum_d);
  YH_LAUNCH_CHECK();
  unsigned long long h[2] = {0, 0};
  YH_CUDA(cudaMemcpyAsync(h, s->sum_d, sizeof(h), cudaMemcpyDeviceToHost, s->main));
  YH_CUDA(cudaStreamSynchronize(s->main));
  *sum_u = h[0]; *sum_v = h[1];
  return YH_OK;
}

int yh_slab_run_host(yh_slab *s, const double *u_in_h, const double *v_in_h, double *u_out_h, double *v_out_h,
                     int nsteps, int tb_steps) {
  YH_REQUIRE(s && u_in_h && v_in_h && u_out_h && v_out_h && nsteps >= 0, "bad arguments");
  YH_REQUIRE(tb_steps == 0 || tb_steps == 1 || tb_steps == 2 || tb_steps == 4, "tb_steps must be 0, 1, 2 or 4");
  bool done = false;
  int rc = run_host_pipelined(&s, 1, &u_in_h, &v_in_h, &u_out_h, &v_out_h, nsteps, tb_steps, &done);
  if (rc != YH_OK) return rc;
  if (done) return yh_slab_sync(s);
  rc = yh_slab_set_state(s, u_in_h, v_in_h, 0);
  if (rc != YH_OK) return rc;
  rc = yh_slab_advance(s, nsteps, tb_steps);
  if (rc != YH_OK) return rc;
  return yh_slab_get_state(s, u_out_h, v_out_h);
}

// Host arithmetic of the pipelined schedule for slab `rank` of `world` slabs of an ny_global-row sheet, no device needed.
static bool geom_from_args(int ny_global, int world, int rank, int halo, int timeIntOrder, int fast, PipeGeom *g) {
  int j0 = 0, j1 = 0;
  if (halo < 1 || timeIntOrder < 1 || yh_slab_partition(ny_global, world, rank, &j0, &j1) != YH_OK) return false;
  const int g0 = j0 - halo > 0 ? j0 - halo : 0;
  g->fast = fast != 0; g->solid = false; g->has_up = rank > 0; g->has_down = rank < world - 1;
  g->halo = halo; g->K = timeIntOrder; g->own_lo = j0 - g0; g->own_hi = j1 - g0; g->ny = ny_global; g->world = world;
  return true;
}

int yh_slab_pipeline_plan(int ny_global, int world, int rank, int halo, int timeIntOrder, int fast, int nsteps,
                          int tb_steps, int plan[8]) {
  PipeGeom g;
  PipePlan pl;
  if (!plan || !geom_from_args(ny_global, world, rank, halo, timeIntOrder, fast, &g)) return 0;
  for (int q = 0; q < 8; q++) plan[q] = 0;
  if (nsteps <= 0 || !plan_geom(g, nsteps, tb_steps, &pl)) return 0;
  plan[0] = pl.n; plan[1] = pl.h; plan[2] = pl.B; plan[3] = pl.P; plan[4] = pl.C; plan[5] = pl.S;
  plan[6] = g.own_lo; plan[7] = g.own_hi;
  return pl.P;
}

int yh_slab_pipeline_region(int ny_global, int world, int rank, int halo, int timeIntOrder, int fast, int nsteps,
                            int tb_steps, int chunk, int level, int rows[2]) {
  PipeGeom g;
  PipePlan pl;
  if (!rows || !geom_from_args(ny_global, world, rank, halo, timeIntOrder, fast, &g)) return YH_ERR_INVALID_ARG;
  if (nsteps <= 0 || !plan_geom(g, nsteps, tb_steps, &pl) || chunk < 0 || chunk >= pl.C || level < 1 || level > pl.P)
    return YH_ERR_INVALID_ARG;
  region_geom(g, pl, chunk, level, &rows[0], &rows[1]);
  return YH_OK;
}

int yh_slab_pipeline_levels(const yh_slab *s, int nsteps, int tb_steps) {
  PipePlan pl;
  return (s && nsteps > 0 && plan_pipeline(s, nsteps, tb_steps, &pl)) ? pl.P : 0;
}

// ---- one process, several devices ----------------------------------------------------------------
int yh_slab_group_create(yh_slab_group **out, const yh_params *pg, int nslabs, const int *devices, int halo) {
  YH_REQUIRE(out && pg && devices && nslabs >= 1, "bad arguments");
  yh_slab_group *g = new yh_slab_group();
  int rc = YH_OK;
  for (int r = 0; r < nslabs && rc == YH_OK; r++) {
    yh_slab *s = nullptr;
    rc = yh_slab_create(&s, pg, r, nslabs, halo, devices[r]);
    if (rc == YH_OK) g->m.push_back(s);
  }
  for (int r = 0; r < nslabs && rc == YH_OK; r++)
    rc = yh_slab_connect_local(g->m[r], r > 0 ? g->m[r - 1] : nullptr, r < nslabs - 1 ? g->m[r + 1] : nullptr);
  if (rc != YH_OK) { yh_slab_group_destroy(g); return rc; }
  *out = g;
  return YH_OK;
}

int yh_slab_group_destroy(yh_slab_group *g) {
  if (!g) return YH_OK;
  for (yh_slab *s : g->m) {          // nobody may still be writing into a block that is about to go
    DevGuard d(s->device);
    cudaStreamSynchronize(s->main); cudaStreamSynchronize(s->edge);
  }
  for (yh_slab *s : g->m) yh_slab_destroy(s);
  delete g;
  return YH_OK;
}

yh_slab *yh_slab_group_member(yh_slab_group *g, int rank) {
  return (g && rank >= 0 && rank < (int)g->m.size()) ? g->m[rank] : nullptr;
}

int yh_slab_group_set_state(yh_slab_group *g, const double *u_h, const double *v_h) {
  YH_REQUIRE(g && u_h && v_h, "null pointer");
  for (yh_slab *s : g->m) {
    const size_t o = (size_t)s->g0 * s->nx;
    int rc = yh_slab_set_state(s, u_h + o, v_h + o, 1);
    if (rc != YH_OK) return rc;
  }
  return YH_OK;
}

int yh_slab_group_get_state(yh_slab_group *g, double *u_h, double *v_h) {
  YH_REQUIRE(g && u_h && v_h, "null pointer");
  for (yh_slab *s : g->m) {          // all copies in flight, then the waits
    DevGuard d(s->device);
    const size_t row = (size_t)s->nx, rows = (size_t)(s->own_hi - s->own_lo), o = (size_t)s->j0 * row;
    YH_CUDA(cudaMemcpyAsync(u_h + o, s->u[s->cur] + (size_t)s->own_lo * row, rows * row * sizeof(double), cudaMemcpyDeviceToHost, s->main));
    YH_CUDA(cudaMemcpyAsync(v_h + o, s->v[s->cur] + (size_t)s->own_lo * row, rows * row * sizeof(double), cudaMemcpyDeviceToHost, s->main));
  }
  return yh_slab_group_sync(g);
}

int yh_slab_group_set_solid(yh_slab_group *g, const uint8_t *mask_h) {
  YH_REQUIRE(g && mask_h, "null pointer");
  for (yh_slab *s : g->m) {
    int rc = yh_slab_set_solid(s, mask_h + (size_t)s->g0 * s->nx);
    if (rc != YH_OK) return rc;
  }
  return YH_OK;
}

int yh_slab_group_advance(yh_slab_group *g, int nsteps, int tb_steps) {
  YH_REQUIRE(g && nsteps >= 0, "bad arguments");
  YH_REQUIRE(tb_steps == 0 || tb_steps == 1 || tb_steps == 2 || tb_steps == 4, "tb_steps must be 0, 1, 2 or 4");
  if (nsteps == 0) return YH_OK;
  for (yh_slab *s : g->m)
    YH_REQUIRE(!s->pg.solidSwitch || s->solid_arg, "solidSwitch set: call yh_slab_group_set_solid first");
  return advance_all(g->m.data(), (int)g->m.size(), nsteps, tb_steps);
}

int yh_slab_group_sync(yh_slab_group *g) {
  YH_REQUIRE(g != nullptr, "null group");
  int rc = YH_OK;
  for (yh_slab *s : g->m) {
    int r = yh_slab_sync(s);
    if (r != YH_OK) rc = r;
  }
  return rc;
}

int yh_slab_group_run_host(yh_slab_group *g, const double *u_in_h, const double *v_in_h, double *u_out_h,
                           double *v_out_h, int nsteps, int tb_steps) {
  YH_REQUIRE(g && u_in_h && v_in_h && u_out_h && v_out_h && nsteps >= 0, "bad arguments");
  YH_REQUIRE(tb_steps == 0 || tb_steps == 1 || tb_steps == 2 || tb_steps == 4, "tb_steps must be 0, 1, 2 or 4");
  {
    const size_t nsl = g->m.size();
    std::vector<const double *> ui(nsl), vi(nsl);
    std::vector<double *> uo(nsl), vo(nsl);
    for (size_t q = 0; q < nsl; q++) {
      const size_t o = (size_t)g->m[q]->j0 * g->m[q]->nx;
      ui[q] = u_in_h + o; vi[q] = v_in_h + o; uo[q] = u_out_h + o; vo[q] = v_out_h + o;
    }
    bool done = false;
    int rc = run_host_pipelined(g->m.data(), (int)nsl, ui.data(), vi.data(), uo.data(), vo.data(), nsteps, tb_steps, &done);
    if (rc != YH_OK) return rc;
    if (done) return yh_slab_group_sync(g);
  }
  int rc = yh_slab_group_set_state(g, u_in_h, v_in_h);
  if (rc != YH_OK) return rc;
  rc = yh_slab_group_advance(g, nsteps, tb_steps);
  if (rc != YH_OK) return rc;
  return yh_slab_group_get_state(g, u_out_h, v_out_h);
}

// ---- symmetry-reduction (co-moving frame) steps on the slabs of a group ("SR mode") ----------
// display()'s reduceSym branch (main.cu:894-954; yh_sim_run_sr on one device) with every kernel on row slabs:
//   exchange ghosts of (u, v)^n                 timeIntOrder + 3 rows per side (RD radius + BFECC radius)
//   RD           rows own-3 .. own+3 -> (u*, v*), velTan                          yh_rd_step
//   tips         owned rows; slabs are contiguous in j, so the LAST tip of the sheet's list (the disc centre)
//                is the last tip of the highest slab that found one                 yh_tip_track_rows
//   integrals    row sums of the owned disc rows; tables added over the slabs (one non-zero contributor per
//                row: exact), closed in the canonical order of the single-sheet pass  yh_sr_integral_rows, _close
//   solve        3x3 on the host (same text as the reference)                       yh_solve_matrix
//   BFECC        owned rows of u* -> (u, v)^{n+1} in the frame moving with c         yh_advect_bfecc_cphi_rows
// The host sees three small results per step (tip counts, the row table, the 12 integrals), as the reference's
// host sees twelve; everything else stays on the devices.  N slabs == yh_sim_run_sr on the whole sheet, bit for
// bit (tests/slab_driver.cu, mode sr).  C++ text of yolohtli_b200/slab.py::SlabRunner._sr_steps.
static int sr_alloc_slab(yh_slab *s) {
  if (s->sr_rows) return YH_OK;
  const size_t n = (size_t)s->ny_local * s->nx * sizeof(double);
  for (int q = 0; q < 2; q++) {
    YH_CUDA(cudaMalloc(&s->sr_vt[q], n));
    YH_CUDA(cudaMalloc(&s->sr_adv[q], n));
    YH_CUDA(cudaMemsetAsync(s->sr_vt[q], 0, n, s->main));     // velTan and the advection field start at zero (main.cu:431-434)
    YH_CUDA(cudaMemsetAsync(s->sr_adv[q], 0, n, s->main));
  }
  s->sr_tip_cap = 65536;
  YH_CUDA(cudaMalloc(&s->sr_tip_count, sizeof(int)));
  YH_CUDA(cudaMalloc(&s->sr_tip_vec, (size_t)s->sr_tip_cap * sizeof(yh_tip)));
  YH_CUDA(cudaMemsetAsync(s->sr_tip_count, 0, sizeof(int), s->main));
  YH_CUDA(cudaMalloc(&s->sr_rows, (size_t)12 * yh_sr_disc_slots(&s->p) * sizeof(double)));
  return YH_OK;
}

int yh_slab_group_sr_state(yh_slab_group *g, double c[3], double phi[3], int set) {
  YH_REQUIRE(g && c && phi, "null pointer");
  for (int q = 0; q < 3; q++) {
    if (set) { g->c[q] = c[q]; g->phi[q] = phi[q]; }
    else { c[q] = g->c[q]; phi[q] = g->phi[q]; }
  }
  if (set) g->sr_count = 0;
  return YH_OK;
}

int yh_slab_group_advance_sr(yh_slab_group *g, int nsteps, double *c_phi_h) {
  YH_REQUIRE(g && nsteps >= 0, "bad arguments");
  const int count = (int)g->m.size();
  int rc;
  for (yh_slab *s : g->m) {
    YH_REQUIRE(s->world == 1 || s->halo >= s->K + 3, "symmetry-reduction steps need timeIntOrder + 3 ghost rows");
    YH_REQUIRE(!s->pg.solidSwitch || s->solid_arg, "solidSwitch set: call yh_slab_group_set_solid first");
    DevGuard d(s->device);
    rc = sr_alloc_slab(s);
    if (rc != YH_OK) return rc;
  }
  const yh_params *pg = &g->m[0]->pg;
  const size_t tbl = (size_t)12 * yh_sr_disc_slots(&g->m[0]->p);
  g->sr_part.resize(tbl); g->sr_sum.resize(tbl);
  double I[12];
  std::vector<int> ntips(count, -1), wave, devs;
  std::vector<yh_tip> last(count);
  const bool dbg = getenv("YH_SR_DEBUG") != nullptr;
  for (int it = 0; it < nsteps; it++) {
    if (dbg) { fprintf(stderr, "sr step %d (count %lld) c = %g %g %g\n", it, g->sr_count, g->c[0], g->c[1], g->c[2]); fflush(stderr); }
    if (count > 1)
      for (yh_slab *s : g->m) {
        DevGuard d(s->device);
        rc = launch_exchange(s, s->cur, s->main);
        if (rc != YH_OK) return rc;
      }
    for (yh_slab *s : g->m) {
      DevGuard d(s->device);
      const int c = s->cur, o = c ^ 1;
      const int r0 = s->own_lo - 3 > 0 ? s->own_lo - 3 : 0, r1 = s->own_hi + 3 < s->ny_local ? s->own_hi + 3 : s->ny_local;
      rc = yh_rd_step(&s->p, s->u[c], s->v[c], s->u[o], s->v[o], s->sr_vt[0], s->sr_vt[1], s->solid_arg, 0,
                      s->nx / 2, s->pg.ny / 2, r0, r1, s->main);
      if (rc != YH_OK) return rc;
    }
    // tips.  Slabs on the SAME device (the one-GPU test vehicle) share that device's look-back scratch of the
    // ordered compaction (yh_workspace), so their tip kernels must not overlap: one slab per device at a time.
    // The last tip of the concatenated list = the disc centre (count == 0 or no tip anywhere: the set centre, B3).
    float cx = (float)pg->tipx0, cy = (float)pg->tipy0;
    {
      ntips.assign(count, -1);
      int left = count;
      while (left > 0) {
        wave.clear(); devs.clear();
        for (int q = 0; q < count; q++) {
          yh_slab *s = g->m[q];
          bool busy = ntips[q] >= 0;
          for (int d : devs) busy = busy || d == s->device;
          if (busy) continue;
          DevGuard d(s->device);
          rc = yh_tip_track_rows(&s->p, s->u[s->cur ^ 1], s->u[s->cur], nullptr, s->sr_tip_count, s->sr_tip_vec, s->sr_tip_cap,
                                 s->p.dt * (double)g->sr_count, s->p.tipAlgorithm, s->own_lo, s->own_hi, s->main);
          if (rc != YH_OK) return rc;
          wave.push_back(q); devs.push_back(s->device);
        }
        for (int q : wave) {
          yh_slab *s = g->m[q];
          DevGuard d(s->device);
          int n = 0;
          YH_CUDA(cudaMemcpyAsync(&n, s->sr_tip_count, sizeof(int), cudaMemcpyDeviceToHost, s->main));
          YH_CUDA(cudaStreamSynchronize(s->main));
          if (n > s->sr_tip_cap) { yh_set_error("tip list overflow on slab %d", s->rank); return YH_ERR_INVALID_ARG; }
          if (n > 0) {
            YH_CUDA(cudaMemcpyAsync(&last[q], s->sr_tip_vec + (n - 1), sizeof(yh_tip), cudaMemcpyDeviceToHost, s->main));
            YH_CUDA(cudaStreamSynchronize(s->main));
          }
          ntips[q] = n;
          left--;
        }
      }
      if (g->sr_count != 0)
        for (int q = 0; q < count; q++)      // rank order = ascending j: the last non-empty list wins
          if (ntips[q] > 0) { cx = last[q].x; cy = last[q].y; }
    }
    if (c_phi_h)
      for (int q = 0; q < 3; q++) { c_phi_h[6 * it + q] = g->c[q]; c_phi_h[6 * it + 3 + q] = g->phi[q]; }   // main.cu:902-903
    const int passes = g->sr_count == 0 ? 2 : 1;                       // first step: main.cu:910-921
    for (int k = 0; k < passes; k++) {
      for (yh_slab *s : g->m) {
        DevGuard d(s->device);
        rc = yh_sr_integral_rows(&s->p, s->u[s->cur], s->v[s->cur], s->sr_vt[0], s->sr_vt[1], s->sr_adv[0], s->sr_adv[1],
                                 cx, cy, s->own_lo, s->own_hi, s->sr_rows, s->main);
        if (rc != YH_OK) return rc;
      }
      for (size_t q = 0; q < tbl; q++) g->sr_sum[q] = 0.0;
      for (yh_slab *s : g->m) {
        DevGuard d(s->device);
        YH_CUDA(cudaMemcpyAsync(g->sr_part.data(), s->sr_rows, tbl * sizeof(double), cudaMemcpyDeviceToHost, s->main));
        YH_CUDA(cudaStreamSynchronize(s->main));
        for (size_t q = 0; q < tbl; q++) g->sr_sum[q] += g->sr_part[q];
      }
      {
        yh_slab *s = g->m[0];
        DevGuard d(s->device);
        if (count > 1) YH_CUDA(cudaMemcpyAsync(s->sr_rows, g->sr_sum.data(), tbl * sizeof(double), cudaMemcpyHostToDevice, s->main));
        rc = yh_sr_integrals_close(&s->p, s->sr_rows, I, s->main);
        if (rc != YH_OK) return rc;
      }
      yh_solve_matrix(g->c, g->phi, I, g->c);                           // main.cu:913 / 926
      if (passes == 2 && k == 0)
        for (yh_slab *s : g->m) {
          DevGuard d(s->device);
          rc = yh_cxy_field(&s->p, s->sr_adv[0], s->sr_adv[1], g->c, g->phi, s->solid_arg, s->main);
          if (rc != YH_OK) return rc;
        }
    }
    for (yh_slab *s : g->m) {
      DevGuard d(s->device);
      const int c = s->cur, o = c ^ 1;
      // u^{n+1} = BFECC(u*) in the frame moving with (c, phi); lands in the `c` buffers (main.cu:930-932)
      rc = yh_advect_bfecc_cphi_rows(&s->p, s->u[o], s->v[o], s->u[c], s->v[c], g->c, g->phi, s->sr_adv[0], s->sr_adv[1],
                                     s->solid_arg, s->own_lo, s->own_hi, s->main);
      if (rc != YH_OK) return rc;
      s->count++; s->raw = false; s->ghosts_valid = false;
    }
    for (int q = 0; q < 3; q++) g->phi[q] = g->phi[q] + g->c[q] * pg->dt;   // main.cu:936-938
    g->sr_count++;
  }
  return yh_slab_group_sync(g);
}

}  // extern "C"
