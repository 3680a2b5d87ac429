// yh_ordered.cuh -- ORDERED atomic compaction shared by the tip tracker (tip.cu) and the contour
// extractor (contour.cu).
//
// The reference appends hits with one atomicAdd per hit (push_back3 / push_back5,
// helper_functions.cu:33-43), so list order depends on block scheduling.  Here CTAs take a ticket
// (atomicAdd), handle the ticket-th run of consecutive cells, and chain their hit counts by
// decoupled look-back, so the list comes out in ascending linear cell index in ONE launch, with
// no sort and no host round trip.  Tickets are handed out in launch order, therefore every CTA a
// look-back waits on is already resident: the spin cannot deadlock.
#pragma once

#include <stdint.h>

struct YhOrdered {
  unsigned long long *state;   // [0] ticket, [1 + b] look-back word of chunk b
  unsigned epoch;              // distinguishes this launch's words from stale ones
  int nchunks;
};

// look-back word: [63:40] epoch, [33:32] flag (1 = aggregate, 2 = inclusive prefix), [31:0] value
__device__ __forceinline__ unsigned long long yh_lb_pack(unsigned epoch, unsigned flag, unsigned v) {
  return ((unsigned long long)(epoch & 0xFFFFFFu) << 40) | ((unsigned long long)flag << 32) | v;
}

// Called by ONE WARP (all 32 lanes, converged) of the CTA that holds `chunk` with the CTA's hit
// total; returns (in every lane) the number of hits in all earlier chunks.  Look-back is
// warp-parallel: lane l examines predecessor chunk-1-l (32 at a time), the nearest one that already
// holds an inclusive prefix ends the walk -- a single-thread walk made the 256-chunk pass of a
// 512^2 sheet a 20 us kernel.  The CTA of the last chunk publishes the grand total to *count
// (replaces cudaMemset(count) + atomicAdd) and re-arms the ticket counter.
__device__ __forceinline__ unsigned yh_ordered_prefix(const YhOrdered &o, int chunk, unsigned total,
                                                      int *count) {
  volatile unsigned long long *st = o.state + 1;
  const int lane = threadIdx.x & 31;
  const unsigned ep = o.epoch & 0xFFFFFFu;
  unsigned prefix = 0;
  if (chunk > 0) {
    if (lane == 0) {
      st[chunk] = yh_lb_pack(o.epoch, 1u, total);
      __threadfence();
    }
    for (int base = chunk - 1;; base -= 32) {
      const int b = base - lane;
      unsigned flag = 2u, val = 0u;   // before the first chunk: an inclusive prefix of zero
      if (b >= 0) {
        unsigned long long w;
        do { w = st[b]; } while ((unsigned)(w >> 40) != ep || ((w >> 32) & 3u) == 0u);
        flag = (unsigned)(w >> 32) & 3u;
        val = (unsigned)w;
      }
      const unsigned incl = __ballot_sync(0xffffffffu, flag == 2u);
      const int first = __ffs(incl) - 1;            // nearest predecessor with an inclusive prefix
      unsigned x = (incl == 0u || lane <= first) ? val : 0u;
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) x += __shfl_xor_sync(0xffffffffu, x, m);
      prefix += x;
      if (incl != 0u) break;
    }
  }
  if (lane == 0) {
    st[chunk] = yh_lb_pack(o.epoch, 2u, prefix + total);
    __threadfence();
    if (chunk == o.nchunks - 1) {
      *count = (int)(prefix + total);
      o.state[0] = 0ull;   // every ticket has been taken
    }
  }
  return prefix;
}

// Block-wide exclusive scan of `mine` in thread order for NT threads; s_warp holds NT/32 ints.
template <int NT>
__device__ __forceinline__ int yh_block_excl_scan(int mine, int *s_warp, int &total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  int woff = 0;
  total = 0;
#pragma unroll
  for (int w = 0; w < NT / 32; w++) {
    if (w < wid) woff += s_warp[w];
    total += s_warp[w];
  }
  return woff + incl - mine;
}
