// rd_quad.cu -- second generation of the temporally blocked Euler + 5-point kernel (rd_fast.cu):
// FOUR columns per thread.  Mode of reactionDiffusion.cu:186-199 + :498-513 (+ :154-184 masks),
// same expression order, bit-identical to T launches of the reference kernel (--fmad=false).
//
// Why: the pair-per-thread kernel is ISSUE bound, not FP64-pipe or HBM bound.  On B200 an FP64
// instruction holds a scheduler's dispatch port for two cycles and nothing else issues in its shadow
// (tools/issue_mix.cu, profiles/r2_issue_mix.txt); ncu of rd_euler_stream<4,128> at 16384^2
// (profiles/r2a_euler_tb4_16384_mix.txt): 64 FP64 + 62 other warp instructions per row pair,
// 2*FP64 + other = 99.4 % of all issue slots.  Only fewer non-FP64 instructions per cell help:
//   * a thread owns a QUAD (4 consecutive cells): loop control, range tests, barriers, addresses
//     are paid once per four cells, the E/W neighbours inside the quad are registers;
//   * rings of THREE rows per level and a loop unrolled by three: every shared-memory access of the
//     levels >= 1 has a compile-time offset from one base register (no (m - c0) & 3 arithmetic);
//   * the W / E neighbours of a row are loaded together with the row (one iteration ahead, kept in
//     registers), so a level reads exactly ONE ring row per iteration and writes one;
//   * the no-flux mirror in x is an ADDRESS (the thread at x = 0 loads column 1 as its W value);
//   * no loader warp: the level-1 warps issue cp.async for their own quads PF rows ahead.
// Shared-memory rows use the 128-byte XOR swizzle (16-byte chunk j of 128-byte line i sits at j ^ (i & 7)),
// so that 32 lanes reading 32-byte quads (stride 32 B) are bank-conflict free.
//
//   level-l row m is produced in iteration  m - c0 + 2l ;  in that iteration the thread loads row
//   m + 1 of level l-1 (written one iteration earlier) as its new N row.
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "rd_euler_cell.cuh"

namespace {

using namespace yh_euler;

struct Quad { double2 a, b; };                       // cells 0,1 | 2,3 of the thread
struct QRow { Quad u, v; double uW, uE, vW, vE; };   // one source row: the quad and its outer neighbours

// shared-memory accesses by 32-bit shared-window address + compile-time offset: the rings are addressed
// as base register + immediate (C++ pointers made ptxas rebuild the window base and add it to every
// offset in every iteration: 13 integer instructions per row)
template <int IMM>
__device__ __forceinline__ double2 lds128(unsigned a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(a), "n"(IMM));
  return v;
}
template <int IMM>
__device__ __forceinline__ double lds64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(IMM));
  return v;
}
template <int IMM>
__device__ __forceinline__ void sts128(unsigned a, const double2 &v) {
  asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(a), "n"(IMM), "d"(v.x), "d"(v.y) : "memory");
}

// byte offset of 16-byte chunk `ch` inside a swizzled row
__device__ __forceinline__ int swz(int ch) { return ((ch >> 3) << 7) | ((((ch & 7) ^ (ch >> 3)) & 7) << 4); }

// FLAV: bit 0 = FAST arithmetic (yh_set_arithmetic), bit 1 = level-0 rows by TMA bulk copies + mbarriers instead of
// per-thread cp.async (YH_EULER_FEED = tma | cpasync; the A/B is in DESIGN.md).
template <int T, int W, bool CANON, bool DEF, bool STIM, bool SOLID, bool FIX, int FLAV>
__global__ void __launch_bounds__(T *(W / 4))
rd_euler_quad(const __grid_constant__ YhK k, const __grid_constant__ FastArgs a) {
  constexpr int H = (T + 1) & ~1;        // halo columns each side
  constexpr int BX = W - 2 * H;          // output columns per strip
  constexpr int NL = W / 4;              // threads per level
  constexpr int FROW = W * 8;            // bytes of one field row (W = 128: 1 KiB = 8 swizzle lines)
  constexpr int ROW = 2 * FROW;          // u then v
  constexpr int NR0 = (T == 1) ? 16 : 8; // level-0 ring (power of two)
  constexpr int PF = NR0 - 3;            // level-0 rows in flight ahead of use: the slot refilled in iteration i was
                                         // last read in iteration i - 2, two barriers earlier
  constexpr int NRL = 3;                 // ring rows of the levels >= 1
  extern __shared__ __align__(1024) unsigned char smraw[];

  const int tid = threadIdx.x;
  const int lev = __shfl_sync(0xffffffffu, tid / NL, 0) + 1;   // warp-uniform (NL is a multiple of 32)
  const int t = tid % NL;
  const int nx = k.nx;
  const int x0 = blockIdx.x * BX, wx0 = x0 - H;
  const int y0 = k.row0 + blockIdx.y * a.RY;
  const int RYe = min(a.RY, k.row1 - y0);
  const int c0 = y0 - T;
  const int dom_lo = -k.jg0, dom_hi = k.nyg - k.jg0;   // local rows that exist globally
  const size_t zoff = (size_t)blockIdx.z * (size_t)a.sim_stride;
  const int n_it = ((RYe + 3 * T + 2) / 3) * 3;

  const int c = 4 * t;                   // window column of the thread's first cell
  const int gx = wx0 + c;
  const bool okA = (gx >= 0) && (gx < nx), okB = (gx + 2 >= 0) && (gx + 2 < nx);
  const bool col_ok = okA || okB;
  const bool outA = okA && (c >= H) && (c + 2 <= W - H);
  const bool outB = okB && (c + 2 >= H) && (c + 4 <= W - H);
  // quads that straddle a domain edge (only when H or nx is not a multiple of 4: FIX variants)
  const bool strL = FIX && (gx == -2), strR = FIX && (gx + 2 == nx);
  const bool canon = CANON && (lev == 1);

  // ---- static shared-memory offsets of this thread (bytes inside a field row) ----------------
  // offA / offB: the thread's own quad in a swizzled row (every store, and the loads of the levels >= 2);
  // srcA / srcB / srcW / srcE: where this level READS its quad and the outer neighbours.  With the TMA feed
  // (FLAV & 2) the level-0 rows are plain row-major copies of the sheet (a bulk copy cannot swizzle a 1-KiB
  // row), so level 1 reads at 32-byte stride -- a 2-way bank conflict on a quarter of the kernel's loads.
  constexpr bool TMA = (FLAV & 2) != 0;
  constexpr int ARITH = FLAV & 1;
  const bool lin = TMA && lev == 1;
  int offA = swz(2 * t), offB = swz(2 * t + 1);
  int srcA = lin ? 32 * t : offA, srcB = lin ? 32 * t + 16 : offB;
  int srcW = (t > 0) ? (lin ? 32 * t - 8 : swz(2 * t - 1) + 8) : srcA;   // column c-1 (no neighbour: any valid address)
  int srcE = (t < NL - 1) ? (lin ? 32 * t + 32 : swz(2 * t + 2)) : srcB;  // column c+4
  if (gx == 0) srcW = srcA + 8;                            // no-flux mirror: W of x = 0 is x = 1
  if (gx + 4 == nx) srcE = srcB;                           //                 E of x = nx-1 is x = nx-2
  // opaque to the compiler from here on: otherwise ptxas rematerialises the swizzle arithmetic from
  // threadIdx inside the row loop instead of keeping the offsets in registers
  asm volatile("" : "+r"(offA), "+r"(offB), "+r"(srcA), "+r"(srcB), "+r"(srcW), "+r"(srcE));

  // rows this level must produce (empty for threads outside the domain)
  const int lo_l = max(dom_lo, y0 - (T - lev));
  const int hi_l = col_ok ? min(dom_hi, y0 + RYe + (T - lev)) : lo_l;
  const int m0 = c0 - 2 * lev;           // row handled in iteration 0
  // iteration windows: rows lo_l-2 .. hi_l-1 are "active" (the first two only fill the registers)
  const int it_first = lo_l - 2 - m0;
  const unsigned n_act = col_ok ? (unsigned)(hi_l - lo_l + 2) : 0u;

  // pacing (batched sweeps): level `lev` performs step count0 + lev - 1 of its sheet
  bool stim_on = false;
  if (STIM) {
    stim_on = k.stim != 0;
    if (a.period) {
      const int per = a.period[blockIdx.z];
      stim_on = per > 0 && ((a.count0 + lev - 1) % per) <= a.duration;
    }
  }

  const unsigned sm_ring0 = (unsigned)__cvta_generic_to_shared(smraw);   // level 0: NR0 rows
  const unsigned src_ring = (lev == 1) ? sm_ring0 : sm_ring0 + NR0 * ROW + (lev - 2) * NRL * ROW;
  const unsigned dst_ring = sm_ring0 + NR0 * ROW + (lev - 1) * NRL * ROW;  // unused by the last level
  // this thread's addresses inside slot 0 of its source / destination ring
  const unsigned aA = src_ring + srcA, aB = src_ring + srcB, aW = src_ring + srcW, aE = src_ring + srcE;
  const unsigned dA = dst_ring + offA, dB = dst_ring + offB;
  // output row pointers of the last level, advanced one row per computed row (no 64-bit multiply per row)
  double *gu = a.u_out + zoff + gx + (ptrdiff_t)lo_l * nx;
  double *gv = a.v_out + zoff + gx + (ptrdiff_t)lo_l * nx;

  // ---- level 0 feed: every level-1 thread fetches its own quad, PF rows ahead ------------------
  const double *__restrict__ u_in = a.u_in + zoff;
  const double *__restrict__ v_in = a.v_in + zoff;
  const int ld_lo = max(dom_lo, c0), ld_hi = min(dom_hi, y0 + RYe + T);
  // running source pointers (row q, column gx).  For a dead pair of an edge quad the address lies up to
  // 16 bytes outside the row; it is never dereferenced (cp.async with src-size 0 reads nothing).
  const double *pu = u_in + gx + (ptrdiff_t)c0 * nx;
  const double *pv = v_in + gx + (ptrdiff_t)c0 * nx;
  int szA = okA ? 16 : 0, szB = okB ? 16 : 0;   // cp.async src-size: 0 = zero fill, nothing read
  asm volatile("" : "+r"(szA), "+r"(szB));      // (kept in registers: ptxas otherwise recomputes okA / okB per row)
  // TMA feed: one thread per CTA moves a whole level-0 row (the strip's columns that exist, u then v) with two
  // bulk copies that complete on the row slot's mbarrier
  const unsigned sm_bar = sm_ring0 + (NR0 + (T - 1) * NRL) * ROW;       // NR0 mbarriers behind the rings
  const int tx0 = max(wx0, 0), tx1 = min(wx0 + W, nx);                   // columns of this strip inside the sheet
  const unsigned t_bytes = (unsigned)(tx1 - tx0) * 8u, t_dst = (unsigned)(tx0 - wx0) * 8u;
  const double *tu = u_in + tx0 + (ptrdiff_t)c0 * nx;
  const double *tv = v_in + tx0 + (ptrdiff_t)c0 * nx;
  if (TMA) {
    if (tid == 0) {
#pragma unroll
      for (int q = 0; q < NR0; q++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sm_bar + 8u * q));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  auto issue_row = [&](int q) {      // called with q = c0, c0+1, ... in order
    if (TMA) {
      if (tid == 0) {
        const unsigned slot = (unsigned)((q - c0) & (NR0 - 1));
        const unsigned dst = sm_ring0 + slot * ROW + t_dst, bar = sm_bar + 8u * slot;
        if (q >= ld_lo && q < ld_hi) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2u * t_bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst), "l"(tu), "r"(t_bytes), "r"(bar) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst + FROW), "l"(tv), "r"(t_bytes), "r"(bar) : "memory");
        } else {   // a row that does not exist (above / below the sheet, past the chunk): the slot's phase still
                   // advances, so that phase parity == use count of the slot for every row
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
        }
      }
      tu += nx; tv += nx;
      return;
    }
    if (q >= ld_lo && q < ld_hi) {
      const unsigned dst = sm_ring0 + (unsigned)((q - c0) & (NR0 - 1)) * ROW;
      cp_async16s(dst + offA, pu, szA);
      cp_async16s(dst + offB, pu + 2, szB);
      cp_async16s(dst + FROW + offA, pv, szA);
      cp_async16s(dst + FROW + offB, pv + 2, szB);
    }
    pu += nx; pv += nx;
    cp_async_commit();
  };
  // wait until level-0 row q has landed (TMA feed: the row slot's mbarrier, phase = use count of the slot)
  auto wait_row = [&](int q) {
    if (q >= ld_lo && q < ld_hi) {
      const unsigned bar = sm_bar + 8u * (unsigned)((q - c0) & (NR0 - 1));
      const unsigned parity = (unsigned)(((q - c0) / NR0) & 1);
      asm volatile("{\n .reg .pred P1;\n LAB_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                   " @P1 bra DONE;\n bra LAB_WAIT;\n DONE:\n }" ::"r"(bar), "r"(parity) : "memory");
    }
  };

  // one source row into registers: base addresses + compile-time offset
  auto ld_row = [&](auto Ic, unsigned rA, unsigned rB, unsigned rW, unsigned rE, QRow &R) {
    constexpr int I = decltype(Ic)::value;
    R.u.a = lds128<I>(rA);
    R.u.b = lds128<I>(rB);
    R.v.a = lds128<I + FROW>(rA);
    R.v.b = lds128<I + FROW>(rB);
    R.uW = lds64<I>(rW);
    R.uE = lds64<I>(rE);
    R.vW = lds64<I + FROW>(rW);
    R.vE = lds64<I + FROW>(rE);
    if (canon) {   // level-0 data is raw: form u0 + (0.0*0.0) literally
      R.u.a.x += 0.0; R.u.a.y += 0.0; R.u.b.x += 0.0; R.u.b.y += 0.0;
      R.v.a.x += 0.0; R.v.a.y += 0.0; R.v.b.x += 0.0; R.v.b.y += 0.0;
      R.uW += 0.0; R.uE += 0.0; R.vW += 0.0; R.vE += 0.0;
    }
    if (FIX) {     // the quad straddles x = 0 or x = nx: put the mirror value in the dead inner cell
      if (strL) { R.u.a.y = R.u.b.y; R.v.a.y = R.v.b.y; }     // x = -1 := x = 1
      if (strR) { R.u.b.x = R.u.a.x; R.v.b.x = R.v.a.x; }     // x = nx := x = nx-2
    }
  };
  // ... at a run-time slot offset (level 1, and the rare rows off the fast path)
  auto ld_row_dyn = [&](int sb, QRow &R) {
    ld_row(std::integral_constant<int, 0>{}, aA + sb, aB + sb, aW + sb, aE + sb, R);
  };

  // mask patterns of the quad, fetched two rows ahead of use (global / L2; 1 B per cell)
  unsigned pat_q0 = 0x1F1F1F1Fu, pat_q1 = 0x1F1F1F1Fu;
  auto ld_pat = [&](int row) -> unsigned {
    const uint8_t *pp = a.pat + (size_t)row * nx + gx;
    const unsigned lo = okA ? *reinterpret_cast<const unsigned short *>(pp) : 0x1F1Fu;
    const unsigned hi = okB ? *reinterpret_cast<const unsigned short *>(pp + 2) : 0x1F1Fu;
    return lo | (hi << 16);
  };

  // "plain" iterations: the row is computed and neither it nor its N row touches a domain edge in y
  const int pl_lo = max(lo_l, dom_lo + 1), pl_hi = min(hi_l, dom_hi - 1);
  const int it_p0 = pl_lo - m0;
  const unsigned n_plain = (col_ok && pl_hi > pl_lo) ? (unsigned)(pl_hi - pl_lo) : 0u;

  // One iteration of one level.  srcN = ring row holding row m+1 of the level below (written one
  // iteration ago), srcM = the ring row that still holds row m-1 (its slot is due for row m+2, which
  // does not exist when m is the last row of the domain), dst = this level's ring row for row m.
  // The no-flux mirrors in y are LOADS from those rows, never register copies: a conditional copy of
  // a whole QRow defeats the renaming of the unrolled S / C / N rotation (ptxas then moves all three
  // rows through registers every iteration -- measured: 125 MOVs per quad row).
  auto row_step = [&](auto Jc, int it, int srcN, int srcM, QRow &S, QRow &C, QRow &N) {
    constexpr int J = decltype(Jc)::value;
    const bool plain = (unsigned)(it - it_p0) < n_plain;
    bool comp = plain;
    if (plain) {
      if (lev == 1) ld_row_dyn(srcN, N);
      else ld_row(std::integral_constant<int, ((J + 2) % NRL) * ROW>{}, aA, aB, aW, aE, N);
    } else if ((unsigned)(it - it_first) < n_act) {   // register fill, first / last row of the domain
      const int m = m0 + it;
      if (m + 1 < dom_hi) {
        if (m + 1 >= dom_lo) ld_row_dyn(srcN, N);
      } else {
        ld_row_dyn(srcM, N);             // last row of the domain: N := row m-1
      }
      if (m >= lo_l) {
        comp = true;
        if (m == dom_lo) ld_row_dyn(srcN, S);   // first row of the domain: S := row m+1
      }
    }
    if (comp) {
      unsigned pat = 0x1F1F1F1Fu;
      if (SOLID) {
        const int m = m0 + it;
        if (m == lo_l) { pat_q0 = ld_pat(m); if (m + 1 < hi_l) pat_q1 = ld_pat(m + 1); }
        pat = pat_q0;
        pat_q0 = pat_q1;
        if (m + 2 < hi_l) pat_q1 = ld_pat(m + 2);
      }
      bool s0 = false, s1 = false, s2 = false, s3 = false;
      if (STIM && stim_on) {
        const int gj = m0 + it + k.jg0;
        s0 = yh_scs_on(k, gx, gj); s1 = yh_scs_on(k, gx + 1, gj);
        s2 = yh_scs_on(k, gx + 2, gj); s3 = yh_scs_on(k, gx + 3, gj);
      }
      Quad uo, vo;
      if (ARITH == 1) {   // FAST flavour (instantiated without stimulus and masks only)
        euler_cell_fast<DEF>(k, a, C.u.a.x, C.v.a.x, C.uW, C.u.a.y, N.u.a.x, S.u.a.x, C.vW, C.v.a.y, N.v.a.x, S.v.a.x, uo.a.x, vo.a.x);
        euler_cell_fast<DEF>(k, a, C.u.a.y, C.v.a.y, C.u.a.x, C.u.b.x, N.u.a.y, S.u.a.y, C.v.a.x, C.v.b.x, N.v.a.y, S.v.a.y, uo.a.y, vo.a.y);
        euler_cell_fast<DEF>(k, a, C.u.b.x, C.v.b.x, C.u.a.y, C.u.b.y, N.u.b.x, S.u.b.x, C.v.a.y, C.v.b.y, N.v.b.x, S.v.b.x, uo.b.x, vo.b.x);
        euler_cell_fast<DEF>(k, a, C.u.b.y, C.v.b.y, C.u.b.x, C.uE, N.u.b.y, S.u.b.y, C.v.b.x, C.vE, N.v.b.y, S.v.b.y, uo.b.y, vo.b.y);
      } else if (!SOLID || __all_sync(__activemask(), pat == 0x1F1F1F1Fu)) {   // all tissue around: plain stencil
        euler_cell<DEF>(k, C.u.a.x, C.v.a.x, C.uW, C.u.a.y, N.u.a.x, S.u.a.x, C.vW, C.v.a.y, N.v.a.x, S.v.a.x, s0, uo.a.x, vo.a.x);
        euler_cell<DEF>(k, C.u.a.y, C.v.a.y, C.u.a.x, C.u.b.x, N.u.a.y, S.u.a.y, C.v.a.x, C.v.b.x, N.v.a.y, S.v.a.y, s1, uo.a.y, vo.a.y);
        euler_cell<DEF>(k, C.u.b.x, C.v.b.x, C.u.a.y, C.u.b.y, N.u.b.x, S.u.b.x, C.v.a.y, C.v.b.y, N.v.b.x, S.v.b.x, s2, uo.b.x, vo.b.x);
        euler_cell<DEF>(k, C.u.b.y, C.v.b.y, C.u.b.x, C.uE, N.u.b.y, S.u.b.y, C.v.b.x, C.vE, N.v.b.y, S.v.b.y, s3, uo.b.y, vo.b.y);
      } else {
        euler_cell_solid<DEF>(k, pat & 0xFFu, C.u.a.x, C.v.a.x, C.uW, C.u.a.y, N.u.a.x, S.u.a.x, C.vW, C.v.a.y, N.v.a.x, S.v.a.x, s0, uo.a.x, vo.a.x);
        euler_cell_solid<DEF>(k, (pat >> 8) & 0xFFu, C.u.a.y, C.v.a.y, C.u.a.x, C.u.b.x, N.u.a.y, S.u.a.y, C.v.a.x, C.v.b.x, N.v.a.y, S.v.a.y, s1, uo.a.y, vo.a.y);
        euler_cell_solid<DEF>(k, (pat >> 16) & 0xFFu, C.u.b.x, C.v.b.x, C.u.a.y, C.u.b.y, N.u.b.x, S.u.b.x, C.v.a.y, C.v.b.y, N.v.b.x, S.v.b.x, s2, uo.b.x, vo.b.x);
        euler_cell_solid<DEF>(k, pat >> 24, C.u.b.y, C.v.b.y, C.u.b.x, C.uE, N.u.b.y, S.u.b.y, C.v.b.x, C.vE, N.v.b.y, S.v.b.y, s3, uo.b.y, vo.b.y);
      }
      if (STIM && a.apd.APD1 && (outA || outB)) {   // fused sAPD epilogue (owner cells)
        const int m = m0 + it;
        if (m >= y0 && m < y0 + RYe) {
          const double th = 0.15;
          const bool e0 = outA && (((C.u.a.x > th) && (uo.a.x < th)) || ((C.u.a.x < th) && (uo.a.x > th)));
          const bool e1 = outA && (((C.u.a.y > th) && (uo.a.y < th)) || ((C.u.a.y < th) && (uo.a.y > th)));
          const bool e2 = outB && (((C.u.b.x > th) && (uo.b.x < th)) || ((C.u.b.x < th) && (uo.b.x > th)));
          const bool e3 = outB && (((C.u.b.y > th) && (uo.b.y < th)) || ((C.u.b.y < th) && (uo.b.y > th)));
          if (e0 || e1 || e2 || e3) {
            const size_t cidx = zoff + (size_t)m * nx + gx;
            if (e0) apd_event(k, a.apd, cidx, uo.a.x, C.u.a.x, a.count0 + lev);
            if (e1) apd_event(k, a.apd, cidx + 1, uo.a.y, C.u.a.y, a.count0 + lev);
            if (e2) apd_event(k, a.apd, cidx + 2, uo.b.x, C.u.b.x, a.count0 + lev);
            if (e3) apd_event(k, a.apd, cidx + 3, uo.b.y, C.u.b.y, a.count0 + lev);
          }
        }
      }
      if (lev < T) {
        sts128<(J % NRL) * ROW>(dA, uo.a);
        sts128<(J % NRL) * ROW>(dB, uo.b);
        sts128<(J % NRL) * ROW + FROW>(dA, vo.a);
        sts128<(J % NRL) * ROW + FROW>(dB, vo.b);
      } else {               // rows are computed in order lo_l, lo_l+1, ...: running pointers
        if (outA) {
          *reinterpret_cast<double2 *>(gu) = uo.a;
          *reinterpret_cast<double2 *>(gv) = vo.a;
        }
        if (outB) {
          *reinterpret_cast<double2 *>(gu + 2) = uo.b;
          *reinterpret_cast<double2 *>(gv + 2) = vo.b;
        }
        gu += nx; gv += nx;
      }
    }
  };

  QRow RA, RB, RC;
  RA.u.a = RA.u.b = RA.v.a = RA.v.b = make_double2(0, 0);
  RA.uW = RA.uE = RA.vW = RA.vE = 0.0;
  RB = RA; RC = RA;

  // One loop for every level (one copy of the row code in the instruction cache).  Iteration `it`
  // reads the row the level below wrote in iteration it-1 and writes slot it % 3 of its own ring.  The
  // loop is unrolled by three, so for the 3-row rings every slot is a compile-time constant; only level
  // 1, whose source is the NR0-row level-0 ring, computes a (warp-uniform) slot.
  if (lev == 1) {
#pragma unroll
    for (int q = 0; q < PF; q++) issue_row(c0 + q);
  }
  auto sub = [&](auto Jc, int it0, QRow &S, QRow &C, QRow &N) {
    constexpr int J = decltype(Jc)::value;
    const int it = it0 + J;
    int sN = ((J + 2) % NRL) * ROW, sM = (J % NRL) * ROW;
    if (lev == 1) {
      issue_row(c0 + PF + it);
      if (TMA) wait_row(c0 + it - 1);    // the row this iteration reads
      else cp_async_wait<PF>();          // this thread's pieces of rows <= c0 + it have landed
      sN = ((it - 1) & (NR0 - 1)) * ROW;
      sM = ((it - 3) & (NR0 - 1)) * ROW;
    }
    __syncthreads();                     // ... everybody's; rows written in the last iteration are visible
    row_step(Jc, it, sN, sM, S, C, N);
  };
  for (int it = 0; it < n_it; it += 3) {
    sub(std::integral_constant<int, 0>{}, it, RA, RB, RC);
    sub(std::integral_constant<int, 1>{}, it, RB, RC, RA);
    sub(std::integral_constant<int, 2>{}, it, RC, RA, RB);
  }
  if (lev == 1 && !TMA) cp_async_wait<0>();
}

template <int T, int W, bool CANON, bool DEF, bool STIM, bool SOLID, bool FIX, int FLAV = 0>
int launch4(const YhK &k, const FastArgs &a, int nsims, cudaStream_t st) {
  constexpr int H = (T + 1) & ~1, BX = W - 2 * H, ROW = 2 * W * 8;
  constexpr int NT = T * (W / 4);
  constexpr int NR0 = (T == 1) ? 16 : 8;
  const size_t smem = (size_t)(NR0 + (T - 1) * 3) * ROW + ((FLAV & 2) ? 8 * NR0 : 0);
  static bool attr_set[64] = {false};
  static int slots[64] = {0};
  int dev = 0;
  YH_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    YH_CUDA(cudaFuncSetAttribute(rd_euler_quad<T, W, CANON, DEF, STIM, SOLID, FIX, FLAV>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1, sms = 148;
    YH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rd_euler_quad<T, W, CANON, DEF, STIM, SOLID, FIX, FLAV>,
                                                          NT, smem));
    YH_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    slots[dev & 63] = (per_sm > 0 ? per_sm : 1) * sms;
    attr_set[dev & 63] = true;
  }
  const int rows = k.row1 - k.row0;
  FastArgs b = a;
  const int strips = (k.nx + BX - 1) / BX;
  b.RY = a.RY > 0 ? a.RY : pick_ry_waves(rows, strips, nsims, T, slots[dev & 63]);
  dim3 grd(strips, (rows + b.RY - 1) / b.RY, nsims);
  auto kfn = rd_euler_quad<T, W, CANON, DEF, STIM, SOLID, FIX, FLAV>;
  YH_LAUNCH(kfn, grd, NT, smem, st, k, b);
  return YH_OK;
}

template <int T, int W, bool CANON, bool DEF, bool FIX>
int launch3(const YhK &k, const FastArgs &a, int nsims, cudaStream_t st) {
  const bool stim = k.stim != 0 || a.period != nullptr || a.apd.APD1 != nullptr;
  if (a.pat)   // obstacle masks: one variant (STIM on) keeps the instantiation count down
    return launch4<T, W, CANON, DEF, true, true, FIX>(k, a, nsims, st);
  if (stim) return launch4<T, W, CANON, DEF, true, false, FIX>(k, a, nsims, st);
  // level-0 feed.  Measured A/B (B200, Gcell/s, cp.async vs TMA bulk copies): T = 1 (HBM-bound) 176 vs 185,
  // T = 2 332 vs 320, T = 4 (issue-bound) 410 vs 379 -- the bulk copy frees the load issue slots of the
  // memory-bound kernel, but its row-major rows cost level 1 a 2-way bank conflict, which the issue-bound
  // kernels feel.  Default: TMA for T = 1, cp.async otherwise; YH_EULER_FEED = tma | cpasync overrides.
  static const int feed = [] { const char *f = getenv("YH_EULER_FEED"); return !f ? -1 : (f[0] == 't' ? 1 : 0); }();
  const bool tma = feed < 0 ? (T == 1) : feed == 1;
  if (yh_arithmetic() == YH_ARITH_FAST) {   // masks and the stimulus stay exact
    FastArgs b = a;
    const double rs = k.gateDiff ? k.rscale : 0.0;
    b.uH = k.tc * k.rx; b.uV = k.tc * k.ry; b.uC = 1.0 - 2.0 * k.tc * (k.rx + k.ry); b.uT = k.tc * k.dt;
    b.vH = rs * b.uH; b.vV = rs * b.uV; b.vC = 1.0 - 2.0 * k.tc * rs * (k.rx + k.ry); b.vT = k.tc * k.dt * k.eps;
    return tma ? launch4<T, W, CANON, DEF, false, false, FIX, 3>(k, b, nsims, st)
               : launch4<T, W, CANON, DEF, false, false, FIX, 1>(k, b, nsims, st);
  }
  return tma ? launch4<T, W, CANON, DEF, false, false, FIX, 2>(k, a, nsims, st)
             : launch4<T, W, CANON, DEF, false, false, FIX>(k, a, nsims, st);
}

template <int T, int W>
int launch(const YhK &k, const FastArgs &a, int nsims, bool canon, cudaStream_t st) {
  constexpr int H = (T + 1) & ~1;
  // DEF variant: every constant whose operation is an exact identity at the reference defaults
  const bool def = (k.tc == 1.0) && (k.mu == 1.0) && (k.delta == 1.0) && (k.gamma == 0.0) && (k.theta == 0.0) &&
                   (k.gateDiff != 0);
  const bool fix = (H % 4 != 0) || (k.nx % 4 != 0);   // a quad can straddle a domain edge
  if (!def) {   // non-default constants: one generic variant per shape
    return canon ? launch3<T, W, true, false, true>(k, a, nsims, st) : launch3<T, W, false, false, true>(k, a, nsims, st);
  }
  if (fix) return canon ? launch3<T, W, true, true, true>(k, a, nsims, st) : launch3<T, W, false, true, true>(k, a, nsims, st);
  return canon ? launch3<T, W, true, true, false>(k, a, nsims, st) : launch3<T, W, false, true, false>(k, a, nsims, st);
}

}  // namespace

int yh_rd_quad_supported(const YhK &k, int tb) {
  (void)tb;
  return k.nx >= 16;
}

// Same contract as yh_launch_rd_fast_paced (rd_fast.cu), which routes here.
int yh_launch_rd_quad_paced(const YhK &k, int tb, const FastArgs &a, int nsims, bool canon, int W,
                            cudaStream_t st) {
#ifdef YH_QUICK   // developer builds: only the bench variant, for SASS inspection
  return launch4<4, 128, false, true, false, false, false, 0>(k, a, nsims, st);
#else
#define YH_QUAD_DISPATCH(WW)                                      \
  switch (tb) {                                                   \
    case 1: return launch<1, WW>(k, a, nsims, canon, st);         \
    case 2: return launch<2, WW>(k, a, nsims, canon, st);         \
    default: return launch<4, WW>(k, a, nsims, canon, st);        \
  }
  if (W == 256) { YH_QUAD_DISPATCH(256) }
  YH_QUAD_DISPATCH(128)
#undef YH_QUAD_DISPATCH
#endif
}
