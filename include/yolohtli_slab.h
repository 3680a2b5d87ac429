/*
 * yolohtli_slab.h -- C ABI of the multi-GPU row-slab driver (libyolohtli_b200.so, csrc/slab.cu).
 *
 * The reference is single-GPU: its host is one C++ loop, display() (main.cu:862-1043), around
 * reactionDiffusion_wrapper + swapSoA (main.cu:879-882).  This is that loop for an nx x ny sheet cut
 * into contiguous row slabs, one slab per GPU: a slab stores its owned rows plus `halo`
 * ghost rows per side, steps its EDGE BANDS first on a high-priority stream, pushes the fresh edge rows
 * straight into the neighbours' ghost rows (NVLink peer stores from a kernel, followed by a release of a
 * sequence number in the neighbour's flag word) while the INTERIOR rows -- which never read a ghost --
 * are still computing on the main stream.  No host involvement on the data path, no NCCL, no Python.
 *
 * Two ways to place slabs:
 *   one process per GPU   yh_slab_create on every rank; yh_slab_export writes an opaque handle (CUDA IPC
 *                         handles of the state block and the flag words); the host moves the handles of
 *                         the two neighbours by whatever it has (MPI_Sendrecv, torch.distributed, a
 *                         file) and passes them to yh_slab_connect.
 *   one process, N GPUs   yh_slab_group_create: creates, peer-enables and connects N slabs; the group
 *                         calls drive all of them from one host thread, like display() drives one.
 *
 * N slabs == one sheet BIT FOR BIT (tests/test_gpu_slab_driver.py, tests/slab_driver.cu).
 * Every entry point returns a status of yolohtli_abi.h (YH_OK == 0).
 */
#ifndef YOLOHTLI_SLAB_H
#define YOLOHTLI_SLAB_H

#include "yolohtli_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct yh_slab yh_slab;
typedef struct yh_slab_group yh_slab_group;

#define YH_SLAB_HANDLE_BYTES 256

/* Rows [*j0, *j1) owned by `rank`: as even as possible, remainder to the low ranks. */
int yh_slab_partition(int ny_global, int world, int rank, int *j0, int *j1);

/* p_global describes the WHOLE sheet (ny == ny_global, jg0 == 0).  halo >= timeIntOrder ghost rows per
 * side (Euler: halo = time steps per exchange, 1 / 2 / 4).  The slab allocates its state on `device`. */
int yh_slab_create(yh_slab **out, const yh_params *p_global, int rank, int world, int halo, int device);
int yh_slab_destroy(yh_slab *s);
/* global row ranges: owned [*j0, *j1), stored (ghosts included) [*g0, *g1) */
int yh_slab_layout(const yh_slab *s, int *j0, int *j1, int *g0, int *g1);

/* ---- wiring ------------------------------------------------------------------------------------ */
int yh_slab_export(yh_slab *s, void *handle /* YH_SLAB_HANDLE_BYTES */);
/* handles of ranks rank-1 / rank+1 made by yh_slab_export in ANOTHER process (NULL at the sheet edges) */
int yh_slab_connect(yh_slab *s, const void *handle_up, const void *handle_down);
/* neighbours living in THIS process (same or peer-accessible device) */
int yh_slab_connect_local(yh_slab *s, yh_slab *up, yh_slab *down);

/* ---- state ------------------------------------------------------------------------------------- */
/* Host rows -> device (asynchronous on the slab's main stream; pinned memory overlaps).  with_ghosts = 0:
 * u_h / v_h hold the owned rows, the ghost rows are fetched from the neighbours by the next advance;
 * with_ghosts = 1: they hold the stored rows [g0, g1). */
int yh_slab_set_state(yh_slab *s, const double *u_h, const double *v_h, int with_ghosts);
/* owned rows -> host; returns after the copy has completed */
int yh_slab_get_state(yh_slab *s, double *u_h, double *v_h);
/* obstacle mask (1 = tissue, main.cu:676-680) of the stored rows [g0, g1) */
int yh_slab_set_solid(yh_slab *s, const uint8_t *mask_h);
void *yh_slab_device_u(yh_slab *s);   /* current state buffers, ny_local x nx, stored rows */
void *yh_slab_device_v(yh_slab *s);
/* the slab's main stream (cudaStream_t): every call above is ordered on it, so events recorded on it
 * bracket the slab's work (bench.py times with CUDA events on this stream) */
void *yh_slab_stream(yh_slab *s);

/* ---- stepping ---------------------------------------------------------------------------------- */
/* nsteps x {reactionDiffusion_wrapper; swapSoA} on this slab, asynchronous.  Every rank must make the
 * same call.  tb_steps: Euler time steps per HBM pass and per exchange (0 = halo). */
int yh_slab_advance(yh_slab *s, int nsteps, int tb_steps);
/* waits for the slab's streams; YH_ERR_CUDA if a neighbour never signalled (flag wait timed out) */
int yh_slab_sync(yh_slab *s);
/* sum of the owned cells' bit patterns (uint64 views) mod 2^64: order independent, so the sum over the
 * ranks is the same number for every decomposition of the same sheet */
int yh_slab_checksum(yh_slab *s, unsigned long long *sum_u, unsigned long long *sum_v);
/* the reference's whole use on one slab with HOST buffers of the owned rows: upload, nsteps, download
 * (main.cu:470 ... 519).  The copies are hidden behind the time steps: the rows travel in chunks, the first
 * blocks of time steps run on a chunk while the next one is in flight, the last blocks while the previous
 * chunk is on its way back (skewed chunk boundaries + edge wedges caught up with one halo exchange per level,
 * csrc/slab.cu); bit-identical to set_state / advance / get_state.  Pinned host memory is what makes the
 * copies asynchronous.  Every rank must make the same call.  YH_SLAB_PIPE=0: plain schedule. */
int yh_slab_run_host(yh_slab *s, const double *u_in_h, const double *v_in_h, double *u_out_h,
                     double *v_out_h, int nsteps, int tb_steps);
/* blocks of time steps per chunk the pipelined schedule of yh_slab_run_host would use for such a call
 * (0: the plain schedule -- run too short, slab too small, masks) */
int yh_slab_pipeline_levels(const yh_slab *s, int nsteps, int tb_steps);
/* The same plan as host arithmetic, no device needed (tests/test_slab_pipeline_plan.py checks its invariants on the
 * CPU): slab `rank` of `world` slabs of an ny_global-row sheet, `halo` ghost rows, fast = the temporally blocked
 * Euler path.  plan[8] = {time steps per block, halo rows per block h, blocks, levels per chunk P, chunks C, chunk
 * height S, first owned LOCAL row, one past the last}; returns P (0: plain schedule).  _region: LOCAL rows
 * [rows[0], rows[1]) of `chunk` at `level` (1 .. P): boundaries move up by h rows per level, an edge that faces a
 * neighbour recedes by h rows per level. */
int yh_slab_pipeline_plan(int ny_global, int world, int rank, int halo, int timeIntOrder, int fast, int nsteps,
                          int tb_steps, int plan[8]);
int yh_slab_pipeline_region(int ny_global, int world, int rank, int halo, int timeIntOrder, int fast, int nsteps,
                            int tb_steps, int chunk, int level, int rows[2]);

/* ---- one process, several devices --------------------------------------------------------------- */
/* devices[r] holds slab r (the same device may appear more than once: test vehicle on one GPU). */
int yh_slab_group_create(yh_slab_group **out, const yh_params *p_global, int nslabs, const int *devices,
                         int halo);
int yh_slab_group_destroy(yh_slab_group *g);
yh_slab *yh_slab_group_member(yh_slab_group *g, int rank);
/* whole-sheet host arrays (nx * ny doubles) */
int yh_slab_group_set_state(yh_slab_group *g, const double *u_h, const double *v_h);
int yh_slab_group_get_state(yh_slab_group *g, double *u_h, double *v_h);
int yh_slab_group_set_solid(yh_slab_group *g, const uint8_t *mask_h);
int yh_slab_group_advance(yh_slab_group *g, int nsteps, int tb_steps);
int yh_slab_group_sync(yh_slab_group *g);
int yh_slab_group_run_host(yh_slab_group *g, const double *u_in_h, const double *v_in_h,
                           double *u_out_h, double *v_out_h, int nsteps, int tb_steps);

/* ---- symmetry-reduction mode on the slabs of a group (display()'s reduceSym branch, main.cu:894-954) ----
 * Every step: ghost exchange, RD + velTan, tips, the 12 phase-condition integrals (row sums per slab, added over
 * the slabs, closed in the single-sheet order), the host 3x3 solve, BFECC in the moving frame -- bit for bit
 * yh_sim_run_sr on the whole sheet.  The group must have been created with halo >= timeIntOrder + 3.
 * c_phi_h (optional): 6 doubles per step, (c, phi) as pushed to clist / philist.  yh_slab_group_sr_state reads
 * or sets (c, phi); setting restarts the step count (the first step solves twice, main.cu:910-921). */
int yh_slab_group_advance_sr(yh_slab_group *g, int nsteps, double *c_phi_h);
int yh_slab_group_sr_state(yh_slab_group *g, double c[3], double phi[3], int set);

#ifdef __cplusplus
}
#endif
#endif /* YOLOHTLI_SLAB_H */
