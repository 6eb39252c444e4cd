// lq_tuned.cuh -- sm_100a-tuned D = 4 kernels of the molecular-dynamics hot loop (CUDA only).
//
// Same arithmetic as the generic functors KEfieldStep / KEfieldLinkStep of lq_kernels.cuh (the parity tests run
// both); what changes is the mapping of threads to links, the register budget and the memory-level parallelism.
#pragma once
#include "lq_kernels.cuh"

#ifndef LQ_HOST_EMU

// MAP: 0 = row walk (lq_site), 1 = tile walk (lq_site_tiled)
template <int MAP>
__device__ __forceinline__ Site<4> lq_tuned_site(const LqGeom& g, lq_i64 n) {
  if (MAP == 1) return lq_site_tiled<4>(g, n);
  return lq_site<4>(g, n);
}

// ------------------------------------------------------------------------------------------------------------
// V1: one thread per link, a warp = 32 consecutive sites of one direction, block = (BLOCK/32) warps covering
// BLOCK/4/32 site groups x 4 directions.  FUSED = 1 also performs the link step into Unew.
template <int BLOCK, int MINB, int MAP, int FUSED>
__global__ void __launch_bounds__(BLOCK, MINB)
    lq_md_link_kernel(LqGeom g, const cx* __restrict__ U, cx* __restrict__ Unew, cx* __restrict__ E, double coef,
                      double dt_e, double dt_u, double c_u, int nkick) {
  constexpr int SITES = BLOCK / 4;
  const int mu = threadIdx.x / SITES;
  const lq_i64 n = (lq_i64)blockIdx.x * SITES + (threadIdx.x - mu * SITES);
  if (n >= g.vol) return;
  const Site<4> st = lq_tuned_site<MAP>(g, n);
  const lq_i64 p = lq_slot<4>(g, st);
  M3 a = lq_staple_sum<4>(U, g, st, mu);
  M3 u = lq_load_link(U, g, mu, p);
  M3 w = m3_mul_nn(u, a);
  cx tr[8];
  lq_trace_gen(w, tr);
  A8 e = lq_load_e(E, g, mu, p);
  for (int kk = 0; kk < nkick; ++kk) {
#pragma unroll
    for (int k = 0; k < 8; ++k) e.e[k] = fma(coef * tr[k].y, dt_e, e.e[k]);
  }
  lq_store_e(E, g, mu, p, e);
  if (FUSED) lq_store_link(Unew, g, mu, p, lq_link_update<4>(u, e, dt_u, c_u, 0));
}

// ------------------------------------------------------------------------------------------------------------
// V2: one thread per (link, nu): the three staple pairs of a link are computed by three warps in parallel and
// summed through shared memory in a fixed order (nu ascending, as the serial loop does), which triples the
// number of independent load streams per link.  Block = 32 sites x 4 mu x 3 nu-slots = 384 threads.
template <int MINB, int MAP, int FUSED>
__global__ void __launch_bounds__(384, MINB)
    lq_md_nusplit_kernel(LqGeom g, const cx* __restrict__ U, cx* __restrict__ Unew, cx* __restrict__ E, double coef,
                         double dt_e, double dt_u, double c_u, int nkick) {
  __shared__ cx sm[8][9][32];  // partial sums of slots 1 and 2, for the 4 directions
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;  // 0..11
  const int mu = w / 3, slot = w - 3 * mu;
  const int nu = slot < mu ? slot : slot + 1;
  const lq_i64 n = (lq_i64)blockIdx.x * 32 + lane;
  const bool live = n < g.vol;
  Site<4> st;
  lq_i64 p = 0;
  M3 acc = m3_zero();
  if (live) {
    st = lq_tuned_site<MAP>(g, n);
    p = lq_slot<4>(g, st);
    const Site<4> xpm = lq_up<4>(g, st, mu);
    {
      const Site<4> xpn = lq_up<4>(g, st, nu);
      M3 a = lq_load_link(U, g, nu, lq_slot<4>(g, xpm));
      M3 b = lq_load_link(U, g, mu, lq_slot<4>(g, xpn));
      M3 t = m3_mul_nd(a, b);
      M3 c = lq_load_link(U, g, nu, p);
      m3_fma_nd(acc, t, c);
    }
    {
      const Site<4> xmn = lq_dn<4>(g, st, nu);
      const Site<4> xpmmn = lq_dn<4>(g, xpm, nu);
      M3 a = lq_load_link(U, g, mu, lq_slot<4>(g, xmn));
      M3 b = lq_load_link(U, g, nu, lq_slot<4>(g, xpmmn));
      M3 t = m3_mul_nn(a, b);
      M3 c = lq_load_link(U, g, nu, lq_slot<4>(g, xmn));
      m3_fma_dn(acc, t, c);
    }
    if (slot > 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) sm[mu * 2 + slot - 1][k][lane] = acc.e[k];
    }
  }
  __syncthreads();
  if (!live || slot != 0) return;
  // fixed summation order: (slot0 + slot1) + slot2  == the serial nu-ascending accumulation up to rounding of
  // the partial sums; deterministic run to run.
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    cx s1 = sm[mu * 2][k][lane], s2 = sm[mu * 2 + 1][k][lane];
    acc.e[k] = cadd(cadd(acc.e[k], s1), s2);
  }
  M3 u = lq_load_link(U, g, mu, p);
  M3 wm = m3_mul_nn(u, acc);
  cx tr[8];
  lq_trace_gen(wm, tr);
  A8 e = lq_load_e(E, g, mu, p);
  for (int kk = 0; kk < nkick; ++kk) {
#pragma unroll
    for (int k = 0; k < 8; ++k) e.e[k] = fma(coef * tr[k].y, dt_e, e.e[k]);
  }
  lq_store_e(E, g, mu, p, e);
  if (FUSED) lq_store_link(Unew, g, mu, p, lq_link_update<4>(u, e, dt_u, c_u, 0));
}

// ------------------------------------------------------------------------------------------------------------
// V3: as V1 but the loop over nu is NOT unrolled (3 iterations, nu = mu+1, mu+2, mu+3 mod 4): a third of the code,
// so the kernel body stays inside the 32 KB instruction cache.
template <int BLOCK, int MINB, int MAP, int FUSED>
__global__ void __launch_bounds__(BLOCK, MINB)
    lq_md_link_loop_kernel(LqGeom g, const cx* __restrict__ U, cx* __restrict__ Unew, cx* __restrict__ E, double coef,
                           double dt_e, double dt_u, double c_u, int nkick) {
  constexpr int SITES = BLOCK / 4;
  const int mu = threadIdx.x / SITES;
  const lq_i64 n = (lq_i64)blockIdx.x * SITES + (threadIdx.x - mu * SITES);
  if (n >= g.vol) return;
  const Site<4> st = lq_tuned_site<MAP>(g, n);
  const lq_i64 p = lq_slot<4>(g, st);
  const Site<4> xpm = lq_up<4>(g, st, mu);
  const lq_i64 ppm = lq_slot<4>(g, xpm);
  M3 acc = m3_zero();
#pragma unroll 1
  for (int j = 1; j < 4; ++j) {
    const int nu = (mu + j) & 3;
    const Site<4> xpn = lq_up<4>(g, st, nu);
    const Site<4> xmn = lq_dn<4>(g, st, nu);
    const Site<4> xpmmn = lq_dn<4>(g, xpm, nu);
    const lq_i64 pmn = lq_slot<4>(g, xmn);
    {
      M3 a = lq_load_link(U, g, nu, ppm);
      M3 b = lq_load_link(U, g, mu, lq_slot<4>(g, xpn));
      M3 t = m3_mul_nd(a, b);
      M3 c = lq_load_link(U, g, nu, p);
      m3_fma_nd(acc, t, c);
    }
    {
      M3 a = lq_load_link(U, g, mu, pmn);
      M3 b = lq_load_link(U, g, nu, lq_slot<4>(g, xpmmn));
      M3 t = m3_mul_nn(a, b);
      M3 c = lq_load_link(U, g, nu, pmn);
      m3_fma_dn(acc, t, c);
    }
  }
  M3 u = lq_load_link(U, g, mu, p);
  M3 w = m3_mul_nn(u, acc);
  cx tr[8];
  lq_trace_gen(w, tr);
  A8 e = lq_load_e(E, g, mu, p);
  for (int kk = 0; kk < nkick; ++kk) {
#pragma unroll
    for (int k = 0; k < 8; ++k) e.e[k] = fma(coef * tr[k].y, dt_e, e.e[k]);
  }
  lq_store_e(E, g, mu, p, e);
  if (FUSED) lq_store_link(Unew, g, mu, p, lq_link_update<4>(u, e, dt_u, c_u, 0));
}

// ------------------------------------------------------------------------------------------------------------
// V4: lean index arithmetic.  One thread per link, a warp = 32 consecutive site slots (one chunk of the chunked-SoA
// layout when ext0 is a multiple of 32) of one direction.  All neighbour slots are p + (sum of per-direction
// deltas): the eight deltas are computed once per thread, every matrix is one 32-bit element index -> IMAD.WIDE ->
// nine LDG.128 with immediate offsets.  The loop over nu is not unrolled (instruction-cache resident body).
__device__ __forceinline__ int lq_sel4(int d, int a0, int a1, int a2, int a3) {
  return d == 0 ? a0 : d == 1 ? a1 : d == 2 ? a2 : a3;
}
__device__ __forceinline__ M3 lq_ld36(const cx* __restrict__ U, int slot, int dir) {
  const int e = ((slot >> 5) * 36 + dir * 9) * 32 + (slot & 31);
  const cx* b = U + e;
  M3 r;
#pragma unroll
  for (int k = 0; k < 9; ++k) r.e[k] = __ldg(b + k * 32);
  return r;
}
// L2 prefetch of one link matrix of the warp (9 planes x the 128-byte lines its slots touch): a hint, two instructions per
// thread.  The lanes of every group of 8 consecutive slots share one line per plane; lane k asks for plane k & 7, all
// lanes for plane 8, so every (plane, line) pair is requested once or more whatever the rotation of lanes inside a row.
template <int L1 = 0>
__device__ __forceinline__ void lq_pf36(const cx* __restrict__ U, int slot, int dir) {
  const int e = ((slot >> 5) * 36 + dir * 9) * 32 + (slot & 31);
  const cx* b = U + e;
  if (L1) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(b + (threadIdx.x & 7) * 32));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(b + 8 * 32));
  } else {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(b + (threadIdx.x & 7) * 32));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(b + 8 * 32));
  }
}
// FLAGS: 32 / 128 = L2 / L1 prefetch of the next (half) stage (both measured slower), 1 = visit nu so that direction 3 (first touched from DRAM by most blocks) comes last, 2 = streaming
// (evict-first) accesses for E and U', 4 = FAKE neighbours (perfect-locality bound, kbench only: wrong results)
template <int BLOCK, int FUSED, int FLAGS, int PUSH>
__device__ __forceinline__ void lq_md4_body(const LqGeom& g, const cx* __restrict__ U, cx* __restrict__ Unew,
                                            cx* __restrict__ E, double coef, double dt_e, double dt_u, double c_u,
                                            int nkick, const LqPush* __restrict__ ps, int blk) {
  constexpr int SITES = BLOCK / 4;
  const int mu = threadIdx.x / SITES;
  const int n = blk * SITES + (threadIdx.x - mu * SITES);
  if (n >= (int)g.vol) return;
  // site decode (row walk, even x0 first)
  const int e0 = g.ext[0], ne0 = g.ne0;
  int row = n / e0;
  const int lane = n - row * e0;
  const int x0 = lane < ne0 ? 2 * lane : 2 * (lane - ne0) + 1;
  int q = row / g.ext[1];
  const int x1 = row - q * g.ext[1] + g.ghost[1];
  row = q;
  q = row / g.ext[2];
  const int x2 = row - q * g.ext[2] + g.ghost[2];
  const int x3 = q + g.ghost[3];
  const int s1 = (int)g.sstride[1], s2 = (int)g.sstride[2], s3 = (int)g.sstride[3];
  const int p = x1 * s1 + x2 * s2 + x3 * s3 + (x0 & 1) * ne0 + (x0 >> 1);
  // slot deltas of the eight neighbours
  const int x0p = x0 + 1 < e0 ? x0 + 1 : 0, x0m = x0 > 0 ? x0 - 1 : e0 - 1;
  const int sl0 = (x0 & 1) * ne0 + (x0 >> 1);
  const int up0 = (x0p & 1) * ne0 + (x0p >> 1) - sl0, dn0 = (x0m & 1) * ne0 + (x0m >> 1) - sl0;
  int up1 = x1 + 1 < g.sext[1] ? s1 : -x1 * s1, dn1 = x1 > 0 ? -s1 : (g.sext[1] - 1) * s1;
  int up2 = x2 + 1 < g.sext[2] ? s2 : -x2 * s2, dn2 = x2 > 0 ? -s2 : (g.sext[2] - 1) * s2;
  int up3 = x3 + 1 < g.sext[3] ? s3 : -x3 * s3, dn3 = x3 > 0 ? -s3 : (g.sext[3] - 1) * s3;
  if (FLAGS & 4) up1 = up2 = up3 = dn1 = dn2 = dn3 = 0;
  if (FLAGS & 8) {
    // L2 prefetch for the blocks one wave ahead: the link chunk that will be their cold (+x3) neighbour row and
    // their own E chunk.  One 128-byte line per thread.
    constexpr int PFD = 640;
    const int nch = (int)g.nchunk;
    int cu = (p >> 5) + (s3 >> 5) + PFD * (SITES / 32);
    cu -= cu >= nch ? nch : 0;
    cu -= cu >= nch ? nch : 0;
    int ce = (p >> 5) + PFD * (SITES / 32);
    ce -= ce >= nch ? nch : 0;
    const char* pu = (const char*)(U + (lq_i64)cu * 36 * 32);
    const char* pe = (const char*)(E + (lq_i64)ce * 16 * 32);
    const int t = threadIdx.x;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(pu + t * 128));
    if (t < 144 - BLOCK) asm volatile("prefetch.global.L2 [%0];" ::"l"(pu + (BLOCK + t) * 128));
    if (t < 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(pe + t * 128));
  }
  const int pm = p + lq_sel4(mu, up0, up1, up2, up3);
  // E early: its latency hides behind the staples
  const int ee = ((p >> 5) * 16 + mu * 4) * 32 + (p & 31);
  cx ev[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) ev[k] = (FLAGS & 2) ? __ldcs(E + ee + k * 32) : E[ee + k * 32];
  M3 acc = m3_zero();
  auto staple_pair = [&](int j) {
    // default: nu = mu+1, mu+2, mu+3 (mod 4); FLAGS&1: nu ascending with the own direction skipped (3 last)
    const int nu = (FLAGS & 1) ? (j - 1 + (j - 1 >= mu ? 1 : 0)) : ((mu + j) & 3);
    const int upn = lq_sel4(nu, up0, up1, up2, up3), dnn = lq_sel4(nu, dn0, dn1, dn2, dn3);
    if ((FLAGS & 32) && j < 3) {  // operands of the next pair: DRAM -> L2 while this pair is computed
      const int n2 = (FLAGS & 1) ? (j + (j >= mu ? 1 : 0)) : ((mu + j + 1) & 3);
      const int up2_ = lq_sel4(n2, up0, up1, up2, up3), dn2_ = lq_sel4(n2, dn0, dn1, dn2, dn3);
      lq_pf36(U, pm, n2);
      lq_pf36(U, p + up2_, mu);
      lq_pf36(U, p, n2);
      lq_pf36(U, p + dn2_, mu);
      lq_pf36(U, pm + dn2_, n2);
      lq_pf36(U, p + dn2_, n2);
    }
    {  // up:  U_nu(x+mu) U_mu^+(x+nu) U_nu^+(x)
      M3 a = lq_ld36(U, pm, nu);
      M3 b = lq_ld36(U, p + upn, mu);
      if (FLAGS & 128) {  // L1 prefetch half a stage ahead: the three operands of the down staple
        lq_pf36<1>(U, p + dnn, mu);
        lq_pf36<1>(U, pm + dnn, nu);
        lq_pf36<1>(U, p + dnn, nu);
      }
      M3 t = m3_mul_nd(a, b);
      M3 c = lq_ld36(U, p, nu);
      m3_fma_nd(acc, t, c);
    }
    {  // down:  (U_mu(x-nu) U_nu(x+mu-nu))^+ U_nu(x-nu)
      if ((FLAGS & 128) && j < 3) {  // ... and of the next up staple
        const int n2 = (FLAGS & 1) ? (j + (j >= mu ? 1 : 0)) : ((mu + j + 1) & 3);
        lq_pf36<1>(U, pm, n2);
        lq_pf36<1>(U, p + lq_sel4(n2, up0, up1, up2, up3), mu);
        lq_pf36<1>(U, p, n2);
      }
      M3 a = lq_ld36(U, p + dnn, mu);
      M3 b = lq_ld36(U, pm + dnn, nu);
      M3 t = m3_mul_nn(a, b);
      M3 c = lq_ld36(U, p + dnn, nu);
      m3_fma_dn(acc, t, c);
    }
  };
  if (FLAGS & 64) {  // fully unrolled: the scheduler may start the loads of the next pair under the current one
    staple_pair(1);
    staple_pair(2);
    staple_pair(3);
  } else {
#pragma unroll 1
    for (int j = 1; j < 4; ++j) staple_pair(j);
  }
  M3 u = lq_ld36(U, p, mu);
  M3 w = m3_mul_nn(u, acc);
  cx tr[8];
  lq_trace_gen(w, tr);
  A8 e;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    e.e[2 * k] = ev[k].x;
    e.e[2 * k + 1] = ev[k].y;
  }
  for (int kk = 0; kk < nkick; ++kk) {
#pragma unroll
    for (int k = 0; k < 8; ++k) e.e[k] = fma(coef * tr[k].y, dt_e, e.e[k]);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (FLAGS & 2) __stcs(E + ee + k * 32, cmk(e.e[2 * k], e.e[2 * k + 1]));
    else E[ee + k * 32] = cmk(e.e[2 * k], e.e[2 * k + 1]);
  }
  if (FUSED) {
    M3 un = lq_link_update<4>(u, e, dt_u, c_u, (FLAGS & 256) ? 1 : 0);  // 256: U <- exp(i dt E) U instead of Euler
    cx* b = Unew + ((p >> 5) * 36 + mu * 9) * 32 + (p & 31);
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      if (FLAGS & 2) __stcs(b + k * 32, un.e[k]);
      else b[k * 32] = un.e[k];
    }
    if (PUSH) {
      const int o2 = g.ghost[2] ? (x2 == 1 ? 0 : (x2 == g.ext[2] ? 2 : 1)) : 1;
      const int o3 = g.ghost[3] ? (x3 == 1 ? 0 : (x3 == g.ext[3] ? 2 : 1)) : 1;
      if (o2 != 1 || o3 != 1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {  // z-face, t-face, zt-corner neighbour
          const int a = c == 1 ? 1 : o2, bb = c == 0 ? 1 : o3;
          if ((a == 1 && bb == 1) || (c == 2 && (o2 == 1 || o3 == 1))) continue;
          const int k = ps->nbmap[a][bb];
          if (k < 0) continue;
          const int pd = p + ps->delta[k];
          cx* d = ps->peer[k] + ((pd >> 5) * 36 + mu * 9) * 32 + (pd & 31);
#pragma unroll
          for (int kk = 0; kk < 9; ++kk) d[kk * 32] = un.e[kk];
        }
      }
    }
  }
}

// FLAGS & 16: persistent walk -- the grid is a few blocks per SM and every block walks a CONTIGUOUS range of rows,
// so the +-x1 neighbour rows of a row were touched by the same SM a moment ago (L1 hits instead of L2 round trips).
template <int BLOCK, int MINB, int FUSED, int FLAGS = 0, int PUSH = 0>
__global__ void __launch_bounds__(BLOCK, MINB)
    lq_md4_kernel(LqGeom g, const cx* __restrict__ U, cx* __restrict__ Unew, cx* __restrict__ E, double coef, double dt_e,
                  double dt_u, double c_u, int nkick, const LqPush* __restrict__ ps, int bps) {
  // ps: device-resident peer table, read by the threads of boundary slices only; bps: blocks per t-slice (0: keep
  // the natural block order)
  if (FLAGS & 16) {
    constexpr int SITES = BLOCK / 4;
    const int nblk = ((int)g.vol + SITES - 1) / SITES;
    const int per = (nblk + (int)gridDim.x - 1) / (int)gridDim.x;
    const int b0 = blockIdx.x * per, b1 = min(b0 + per, nblk);
#pragma unroll 1
    for (int b = b0; b < b1; ++b) lq_md4_body<BLOCK, FUSED, FLAGS, 0>(g, U, Unew, E, coef, dt_e, dt_u, c_u, nkick, ps, b);
    return;
  }
  int blk = blockIdx.x;
  if (PUSH && bps) {
    // The blocks of the two boundary t-slices are interleaved 1 : (S-1) with interior blocks over the first part of
    // the grid: their NVLink stores are spread over S times their own compute time instead of saturating the link
    // in one burst, and everything has landed long before the kernel ends.  (ext3 < 2S: first, last, interior.)
    constexpr int S = 4;
    const int nbb = 2 * bps;
    if (g.ext[3] >= 2 * S) {
      const int j = blk / S;
      if (blk - j * S == 0 && j < nbb) {
        blk = j < bps ? j : (g.ext[3] - 1) * bps + (j - bps);
      } else {
        const int before = min((blk + S - 1) / S, nbb);
        blk = bps + (blk - before);
      }
    } else {
      const int sl = blk / bps, r = blk - sl * bps;
      blk = (sl == 0 ? 0 : sl == 1 ? g.ext[3] - 1 : sl - 1) * bps + r;
    }
  }
  lq_md4_body<BLOCK, FUSED, FLAGS, PUSH>(g, U, Unew, E, coef, dt_e, dt_u, c_u, nkick, ps, blk);
}

// ------------------------------------------------------------------------------------------------------------
// V5: V4 with the loads software-pipelined by hand.  The staple sum is a stream of six (A, B, C) triples; the
// loads of the next triple are issued before the two matrix products of the current one, so a warp hides its own
// L2/DRAM latency behind ~430 DFMAs instead of relying on the two other warps of its scheduler (ncu of V4: 39 %
// of the stall samples are long-scoreboard waits on the first DFMA that touches a freshly loaded matrix).
// PIPE: 1 = prefetch A,B of the next half-stage; 2 = also fence the order with compiler barriers.
#define LQ_CBAR() asm volatile("" ::: "memory")
template <int BLOCK, int MINB, int FUSED, int PIPE = 1>
__global__ void __launch_bounds__(BLOCK, MINB)
    lq_md5_kernel(LqGeom g, const cx* __restrict__ U, cx* __restrict__ Unew, cx* __restrict__ E, double coef, double dt_e,
                  double dt_u, double c_u, int nkick) {
  constexpr int SITES = BLOCK / 4;
  const int mu = threadIdx.x / SITES;
  const int n = blockIdx.x * SITES + (threadIdx.x - mu * SITES);
  if (n >= (int)g.vol) return;
  const int e0 = g.ext[0], ne0 = g.ne0;
  int row = n / e0;
  const int lane = n - row * e0;
  const int x0 = lane < ne0 ? 2 * lane : 2 * (lane - ne0) + 1;
  int q = row / g.ext[1];
  const int x1 = row - q * g.ext[1] + g.ghost[1];
  row = q;
  q = row / g.ext[2];
  const int x2 = row - q * g.ext[2] + g.ghost[2];
  const int x3 = q + g.ghost[3];
  const int s1 = (int)g.sstride[1], s2 = (int)g.sstride[2], s3 = (int)g.sstride[3];
  const int p = x1 * s1 + x2 * s2 + x3 * s3 + (x0 & 1) * ne0 + (x0 >> 1);
  const int x0p = x0 + 1 < e0 ? x0 + 1 : 0, x0m = x0 > 0 ? x0 - 1 : e0 - 1;
  const int sl0 = (x0 & 1) * ne0 + (x0 >> 1);
  const int up0 = (x0p & 1) * ne0 + (x0p >> 1) - sl0, dn0 = (x0m & 1) * ne0 + (x0m >> 1) - sl0;
  const int up1 = x1 + 1 < g.sext[1] ? s1 : -x1 * s1, dn1 = x1 > 0 ? -s1 : (g.sext[1] - 1) * s1;
  const int up2 = x2 + 1 < g.sext[2] ? s2 : -x2 * s2, dn2 = x2 > 0 ? -s2 : (g.sext[2] - 1) * s2;
  const int up3 = x3 + 1 < g.sext[3] ? s3 : -x3 * s3, dn3 = x3 > 0 ? -s3 : (g.sext[3] - 1) * s3;
  const int pm = p + lq_sel4(mu, up0, up1, up2, up3);
  const int ee = ((p >> 5) * 16 + mu * 4) * 32 + (p & 31);
  M3 acc = m3_zero();
  // prologue: A, B of the first up-staple
  int nu = (mu + 1) & 3;
  int upn = lq_sel4(nu, up0, up1, up2, up3), dnn = lq_sel4(nu, dn0, dn1, dn2, dn3);
  M3 a = lq_ld36(U, pm, nu);
  M3 b = lq_ld36(U, p + upn, mu);
  cx ev[4];
#pragma unroll 1
  for (int j = 1; j < 4; ++j) {
    M3 c = lq_ld36(U, p, nu);
    M3 ad = lq_ld36(U, p + dnn, mu);
    M3 bd = lq_ld36(U, pm + dnn, nu);
    if (PIPE & 2) LQ_CBAR();
    {  // up:  U_nu(x+mu) U_mu^+(x+nu) U_nu^+(x)
      M3 t = m3_mul_nd(a, b);
      m3_fma_nd(acc, t, c);
    }
    if (PIPE & 2) LQ_CBAR();
    c = lq_ld36(U, p + dnn, nu);
    if (j < 3) {
      nu = (mu + j + 1) & 3;
      upn = lq_sel4(nu, up0, up1, up2, up3);
      dnn = lq_sel4(nu, dn0, dn1, dn2, dn3);
      a = lq_ld36(U, pm, nu);
      b = lq_ld36(U, p + upn, mu);
    } else {
      a = lq_ld36(U, p, mu);  // the link itself, for U * A
#pragma unroll
      for (int k = 0; k < 4; ++k) ev[k] = __ldcs(E + ee + k * 32);
    }
    if (PIPE & 2) LQ_CBAR();
    {  // down:  (U_mu(x-nu) U_nu(x+mu-nu))^+ U_nu(x-nu)
      M3 t = m3_mul_nn(ad, bd);
      m3_fma_dn(acc, t, c);
    }
  }
  const M3 u = a;
  M3 w = m3_mul_nn(u, acc);
  cx tr[8];
  lq_trace_gen(w, tr);
  A8 e;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    e.e[2 * k] = ev[k].x;
    e.e[2 * k + 1] = ev[k].y;
  }
  for (int kk = 0; kk < nkick; ++kk) {
#pragma unroll
    for (int k = 0; k < 8; ++k) e.e[k] = fma(coef * tr[k].y, dt_e, e.e[k]);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) __stcs(E + ee + k * 32, cmk(e.e[2 * k], e.e[2 * k + 1]));
  if (FUSED) {
    M3 un = lq_link_update<4>(u, e, dt_u, c_u, 0);
    cx* bo = Unew + ((p >> 5) * 36 + mu * 9) * 32 + (p & 31);
#pragma unroll
    for (int k = 0; k < 9; ++k) __stcs(bo + k * 32, un.e[k]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Checkerboard sweep sub-step (one direction mu, one colour), D = 4, with the lean addressing of V4: the staple sum
// of KHeatBath / KOverrelax (same accumulation order: nu ascending, up then down => the same bits as the generic
// functors) followed by the single-link rule.  KIND: 0 heat bath (heat_bath.rs:73-123), 1 over-relaxation
// (overrelaxation.rs:86-110, 158-184).  Links of the updated (mu, colour) set never enter each other's staples, so
// neighbours are read through the non-coherent path while the own link is read and written in place.
// PF: 0 none; 1 = L2 prefetch of the operands of the next nu pair; 2 = of all 19 matrices at thread start
template <int BLOCK, int MINB, int KIND, int PF = 0>
__global__ void __launch_bounds__(BLOCK, MINB)
    lq_sweep4_kernel(LqGeom g, cx* __restrict__ U, int mu, int parity, int flags, int or_kind, double coupling,
                     unsigned long long seed, unsigned long long counter) {
  const int n = blockIdx.x * BLOCK + threadIdx.x;
  if (n >= (int)(g.vol >> 1)) return;
  const int e0 = g.ext[0], ne0 = g.ne0, h0 = e0 >> 1;
  int row = n / h0;
  const int k = n - row * h0;
  int q = row / g.ext[1];
  const int i1 = row - q * g.ext[1];
  row = q;
  q = row / g.ext[2];
  const int i2 = row - q * g.ext[2];
  const int i3 = q;
  const int x0 = 2 * k + ((parity + i1 + g.goff[1] + i2 + g.goff[2] + i3 + g.goff[3] + g.goff[0]) & 1);
  const int x1 = i1 + g.ghost[1], x2 = i2 + g.ghost[2], x3 = i3 + g.ghost[3];
  const int s1 = (int)g.sstride[1], s2 = (int)g.sstride[2], s3 = (int)g.sstride[3];
  const int sl0 = (x0 & 1) * ne0 + (x0 >> 1);
  const int p = x1 * s1 + x2 * s2 + x3 * s3 + sl0;
  const int x0p = x0 + 1 < e0 ? x0 + 1 : 0, x0m = x0 > 0 ? x0 - 1 : e0 - 1;
  const int up0 = (x0p & 1) * ne0 + (x0p >> 1) - sl0, dn0 = (x0m & 1) * ne0 + (x0m >> 1) - sl0;
  const int up1 = x1 + 1 < g.sext[1] ? s1 : -x1 * s1, dn1 = x1 > 0 ? -s1 : (g.sext[1] - 1) * s1;
  const int up2 = x2 + 1 < g.sext[2] ? s2 : -x2 * s2, dn2 = x2 > 0 ? -s2 : (g.sext[2] - 1) * s2;
  const int up3 = x3 + 1 < g.sext[3] ? s3 : -x3 * s3, dn3 = x3 > 0 ? -s3 : (g.sext[3] - 1) * s3;
  const int pm = p + lq_sel4(mu, up0, up1, up2, up3);
  M3 acc = m3_zero();
  auto pf_pair = [&](int j2) {
    const int n2 = j2 + (j2 >= mu ? 1 : 0);
    const int u2 = lq_sel4(n2, up0, up1, up2, up3), d2 = lq_sel4(n2, dn0, dn1, dn2, dn3);
    lq_pf36(U, pm, n2);
    lq_pf36(U, p + u2, mu);
    lq_pf36(U, p, n2);
    lq_pf36(U, p + d2, mu);
    lq_pf36(U, pm + d2, n2);
    lq_pf36(U, p + d2, n2);
  };
  if (PF == 2) {
    pf_pair(1);
    pf_pair(2);
    lq_pf36(U, p, mu);
  }
#pragma unroll 1
  for (int j = 0; j < 3; ++j) {
    const int nu = j + (j >= mu ? 1 : 0);  // ascending, own direction skipped
    const int upn = lq_sel4(nu, up0, up1, up2, up3), dnn = lq_sel4(nu, dn0, dn1, dn2, dn3);
    if (PF == 1) {
      if (j < 2) pf_pair(j + 1);
      else lq_pf36(U, p, mu);
    }
    {
      M3 a = lq_ld36(U, pm, nu);
      M3 b = lq_ld36(U, p + upn, mu);
      M3 t = m3_mul_nd(a, b);
      M3 c = lq_ld36(U, p, nu);
      m3_fma_nd(acc, t, c);
    }
    {
      M3 a = lq_ld36(U, p + dnn, mu);
      M3 b = lq_ld36(U, pm + dnn, nu);
      M3 t = m3_mul_nn(a, b);
      M3 c = lq_ld36(U, p + dnn, nu);
      m3_fma_dn(acc, t, c);
    }
  }
  cx* own = U + ((p >> 5) * 36 + mu * 9) * 32 + (p & 31);
  M3 u;
#pragma unroll
  for (int kk = 0; kk < 9; ++kk) u.e[kk] = own[kk * 32];
  M3 r;
  if (KIND == 0) {
    // global reference-order link index = RNG stream id (results do not depend on the decomposition)
    const lq_i64 gi = (lq_i64)(x0 + g.goff[0]) * g.gstride[0] + (lq_i64)(i1 + g.goff[1]) * g.gstride[1] +
                      (lq_i64)(i2 + g.goff[2]) * g.gstride[2] + (lq_i64)(i3 + g.goff[3]) * g.gstride[3];
    LqStream rng(seed, counter, (uint64_t)(gi * 4 + mu));
    r = lq_heat_bath_link(u, acc, coupling, rng, flags);
  } else {
    r = lq_overrelax_link(u, acc, or_kind);
  }
#pragma unroll
  for (int kk = 0; kk < 9; ++kk) own[kk * 32] = r.e[kk];
}

// ------------------------------------------------------------------------------------------------------------
// Gauss projection step, D = 4 (project_to_gauss_step, field.rs:1301-1337): the arithmetic of KGaussProjectStep
// (lq_gauss_project_link) with the lean 32-bit addressing of V4 -- one thread per link, a warp = 32 consecutive site slots
// of one direction, read-only loads of U and G, streaming accesses of E (0.206 vs 0.2175 ms for the generic functor,
// 0.216 vs 0.237 ms inside a trajectory at 32^4).
template <int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
    lq_gstep4_kernel(LqGeom g, const cx* __restrict__ U, const cx* __restrict__ G, const cx* __restrict__ Ein,
                     cx* __restrict__ Eout) {
  constexpr int SITES = BLOCK / 4;
  const int mu = threadIdx.x / SITES;
  const int n = blockIdx.x * SITES + (threadIdx.x - mu * SITES);
  if (n >= (int)g.vol) return;
  const int e0 = g.ext[0], ne0 = g.ne0;
  int row = n / e0;
  const int lane = n - row * e0;
  const int x0 = lane < ne0 ? 2 * lane : 2 * (lane - ne0) + 1;
  int q = row / g.ext[1];
  const int x1 = row - q * g.ext[1] + g.ghost[1];
  row = q;
  q = row / g.ext[2];
  const int x2 = row - q * g.ext[2] + g.ghost[2];
  const int x3 = q + g.ghost[3];
  const int s1 = (int)g.sstride[1], s2 = (int)g.sstride[2], s3 = (int)g.sstride[3];
  const int sl0 = (x0 & 1) * ne0 + (x0 >> 1);
  const int p = x1 * s1 + x2 * s2 + x3 * s3 + sl0;
  int up;
  if (mu == 0) {
    const int x0p = x0 + 1 < e0 ? x0 + 1 : 0;
    up = (x0p & 1) * ne0 + (x0p >> 1) - sl0;
  } else if (mu == 1) {
    up = x1 + 1 < g.sext[1] ? s1 : -x1 * s1;
  } else if (mu == 2) {
    up = x2 + 1 < g.sext[2] ? s2 : -x2 * s2;
  } else {
    up = x3 + 1 < g.sext[3] ? s3 : -x3 * s3;
  }
  const cx* gb = G + ((p >> 5) * 9) * 32 + (p & 31);
  const int pp = p + up;
  const cx* gpb = G + ((pp >> 5) * 9) * 32 + (pp & 31);
  const int ee = ((p >> 5) * 16 + mu * 4) * 32 + (p & 31);
  A8 e;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const cx v = __ldcs(Ein + ee + k * 32);
    e.e[2 * k] = v.x;
    e.e[2 * k + 1] = v.y;
  }
  const M3 u = lq_ld36(U, p, mu);
  M3 gx, gp;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    gx.e[k] = __ldg(gb + k * 32);
    gp.e[k] = __ldg(gpb + k * 32);
  }
  e = lq_gauss_project_link(u, gx, gp, e);
#pragma unroll
  for (int k = 0; k < 4; ++k) __stcs(Eout + ee + k * 32, cmk(e.e[2 * k], e.e[2 * k + 1]));
}

// ------------------------------------------------------------------------------------------------------------
// Gauss field, D = 4 (EField::gauss, field.rs:1174-1195): the arithmetic and summation order of lq_gauss_site with the
// lean 32-bit addressing; one thread per site.  PUSH: boundary sites also go straight into the neighbour ranks' ghost
// layers (peer memory), as in KGaussField.
template <int BLOCK, int MINB, int PUSH>
__global__ void __launch_bounds__(BLOCK, MINB)
    lq_gfield4_kernel(LqGeom g, const cx* __restrict__ U, const cx* __restrict__ E, cx* __restrict__ G,
                      const LqPush* __restrict__ ps) {
  const int n = blockIdx.x * BLOCK + threadIdx.x;
  if (n >= (int)g.vol) return;
  const int e0 = g.ext[0], ne0 = g.ne0;
  int row = n / e0;
  const int lane = n - row * e0;
  const int x0 = lane < ne0 ? 2 * lane : 2 * (lane - ne0) + 1;
  int q = row / g.ext[1];
  const int x1 = row - q * g.ext[1] + g.ghost[1];
  row = q;
  q = row / g.ext[2];
  const int x2 = row - q * g.ext[2] + g.ghost[2];
  const int x3 = q + g.ghost[3];
  const int s1 = (int)g.sstride[1], s2 = (int)g.sstride[2], s3 = (int)g.sstride[3];
  const int sl0 = (x0 & 1) * ne0 + (x0 >> 1);
  const int p = x1 * s1 + x2 * s2 + x3 * s3 + sl0;
  const int x0m = x0 > 0 ? x0 - 1 : e0 - 1;
  const int dn[4] = {(x0m & 1) * ne0 + (x0m >> 1) - sl0, x1 > 0 ? -s1 : (g.sext[1] - 1) * s1,
                     x2 > 0 ? -s2 : (g.sext[2] - 1) * s2, x3 > 0 ? -s3 : (g.sext[3] - 1) * s3};
  M3 acc = m3_zero();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pm = p + dn[i];
    A8 eo, em;
    const cx* eb = E + ((p >> 5) * 16 + i * 4) * 32 + (p & 31);
    const cx* mb = E + ((pm >> 5) * 16 + i * 4) * 32 + (pm & 31);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const cx a = __ldg(eb + k * 32), b = __ldg(mb + k * 32);
      eo.e[2 * k] = a.x;
      eo.e[2 * k + 1] = a.y;
      em.e[2 * k] = b.x;
      em.e[2 * k + 1] = b.y;
    }
    acc = m3_add(acc, lq_adjoint_to_matrix(eo));
    const M3 u = lq_ld36(U, pm, i);
    const M3 t = m3_mul_dn(u, lq_adjoint_to_matrix(em));  // U^+ E
    M3 neg = m3_zero();
    m3_fma_nn(neg, t, u);
    acc = m3_sub(acc, neg);
  }
  cx* gb = G + ((p >> 5) * 9) * 32 + (p & 31);
#pragma unroll
  for (int k = 0; k < 9; ++k) gb[k * 32] = acc.e[k];
  if (PUSH) {
    const int o2 = g.ghost[2] ? (x2 == 1 ? 0 : (x2 == g.ext[2] ? 2 : 1)) : 1;
    const int o3 = g.ghost[3] ? (x3 == 1 ? 0 : (x3 == g.ext[3] ? 2 : 1)) : 1;
    if (o2 != 1 || o3 != 1) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {  // z-face, t-face, zt-corner neighbour
        const int a = c == 1 ? 1 : o2, bb = c == 0 ? 1 : o3;
        if ((a == 1 && bb == 1) || (c == 2 && (o2 == 1 || o3 == 1))) continue;
        const int nb = ps->nbmap[a][bb];
        if (nb < 0) continue;
        const int pd = p + ps->delta[nb];
        cx* d = ps->peer[nb] + ((pd >> 5) * 9) * 32 + (pd & 31);
#pragma unroll
        for (int k = 0; k < 9; ++k) d[k * 32] = acc.e[k];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Metropolis sub-step (one direction, one colour), D = 4: MetropolisHastingsSweep (metropolis_hastings_sweep.rs:126-174)
// with the lean addressing of lq_sweep4_kernel and the same arithmetic as KMetropolis (staple order nu ascending, up
// then down; proposal draws, then the accept draw, from the link's Philox stream).  The block's (#accepted, sum of
// acceptance probabilities) go to partial[2 * blockIdx]: the eight sub-steps of a sweep write eight consecutive
// segments and ONE final reduction (and one host synchronisation) serves the whole sweep.
template <int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
    lq_metro4_kernel(LqGeom g, cx* __restrict__ U, int mu, int parity, int flags, int n_update, double beta, double CA,
                     double spread, unsigned long long seed, unsigned long long counter, double* __restrict__ partial) {
  const int n = blockIdx.x * BLOCK + threadIdx.x;
  double v0 = 0.0, v1 = 0.0;
  if (n < (int)(g.vol >> 1)) {
    const int e0 = g.ext[0], ne0 = g.ne0, h0 = e0 >> 1;
    int row = n / h0;
    const int k = n - row * h0;
    int q = row / g.ext[1];
    const int i1 = row - q * g.ext[1];
    row = q;
    q = row / g.ext[2];
    const int i2 = row - q * g.ext[2];
    const int i3 = q;
    const int x0 = 2 * k + ((parity + i1 + g.goff[1] + i2 + g.goff[2] + i3 + g.goff[3] + g.goff[0]) & 1);
    const int x1 = i1 + g.ghost[1], x2 = i2 + g.ghost[2], x3 = i3 + g.ghost[3];
    const int s1 = (int)g.sstride[1], s2 = (int)g.sstride[2], s3 = (int)g.sstride[3];
    const int sl0 = (x0 & 1) * ne0 + (x0 >> 1);
    const int p = x1 * s1 + x2 * s2 + x3 * s3 + sl0;
    const int x0p = x0 + 1 < e0 ? x0 + 1 : 0, x0m = x0 > 0 ? x0 - 1 : e0 - 1;
    const int up0 = (x0p & 1) * ne0 + (x0p >> 1) - sl0, dn0 = (x0m & 1) * ne0 + (x0m >> 1) - sl0;
    const int up1 = x1 + 1 < g.sext[1] ? s1 : -x1 * s1, dn1 = x1 > 0 ? -s1 : (g.sext[1] - 1) * s1;
    const int up2 = x2 + 1 < g.sext[2] ? s2 : -x2 * s2, dn2 = x2 > 0 ? -s2 : (g.sext[2] - 1) * s2;
    const int up3 = x3 + 1 < g.sext[3] ? s3 : -x3 * s3, dn3 = x3 > 0 ? -s3 : (g.sext[3] - 1) * s3;
    const int pm = p + lq_sel4(mu, up0, up1, up2, up3);
    M3 acc = m3_zero();
#pragma unroll 1
    for (int j = 0; j < 3; ++j) {
      const int nu = j + (j >= mu ? 1 : 0);  // ascending, own direction skipped
      const int upn = lq_sel4(nu, up0, up1, up2, up3), dnn = lq_sel4(nu, dn0, dn1, dn2, dn3);
      {
        M3 a = lq_ld36(U, pm, nu);
        M3 b = lq_ld36(U, p + upn, mu);
        M3 t = m3_mul_nd(a, b);
        M3 c = lq_ld36(U, p, nu);
        m3_fma_nd(acc, t, c);
      }
      {
        M3 a = lq_ld36(U, p + dnn, mu);
        M3 b = lq_ld36(U, pm + dnn, nu);
        M3 t = m3_mul_nn(a, b);
        M3 c = lq_ld36(U, p + dnn, nu);
        m3_fma_dn(acc, t, c);
      }
    }
    cx* own = U + ((p >> 5) * 36 + mu * 9) * 32 + (p & 31);
    M3 old;
#pragma unroll
    for (int kk = 0; kk < 9; ++kk) old.e[kk] = own[kk * 32];
    const lq_i64 gi = (lq_i64)(x0 + g.goff[0]) * g.gstride[0] + (lq_i64)(i1 + g.goff[1]) * g.gstride[1] +
                      (lq_i64)(i2 + g.goff[2]) * g.gstride[2] + (lq_i64)(i3 + g.goff[3]) * g.gstride[3];
    LqStream rng(seed, counter, (uint64_t)(gi * 4 + mu));
    const M3 prop = lq_metropolis_proposal(old, n_update, spread, rng, flags);
    const double proba = fmax(fmin(exp(-lq_delta_s(acc, prop, old, beta, CA)), 1.0), 0.0);
    v1 = proba;
    if (rng.bernoulli(proba)) {
      v0 = 1.0;
#pragma unroll
      for (int kk = 0; kk < 9; ++kk) own[kk * 32] = prop.e[kk];
    }
  }
  __shared__ double sm[2][BLOCK / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v0 += __shfl_down_sync(0xffffffffu, v0, o);
    v1 += __shfl_down_sync(0xffffffffu, v1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sm[0][threadIdx.x >> 5] = v0;
    sm[1][threadIdx.x >> 5] = v1;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double x = 0.0;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; ++w) x += sm[threadIdx.x][w];
    partial[(lq_i64)blockIdx.x * 2 + threadIdx.x] = x;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Plaquette reduction, D = 4 (average_trace_plaquette field.rs:775-804, hamiltonian_links state.rs:821-849): the terms
// of KPlaquette with the lean addressing of V4.  One thread per (site, plane i < j), a warp = 32 consecutive site slots
// of one plane, a block = 32 sites x 6 planes.  Per-block partial sums (sum Re Tr P, sum Im Tr P, sum (1 - Re Tr P/CA))
// in a fixed order: warp shuffle tree, then the six warps in plane order; lq_final_k<3> adds the blocks.
template <int MINB>
__global__ void __launch_bounds__(192, MINB)
    lq_plaq4_kernel(LqGeom g, const cx* __restrict__ U, double CA, double* __restrict__ partial) {
  const int pl = threadIdx.x >> 5, lane32 = threadIdx.x & 31;
  const int n = blockIdx.x * 32 + lane32;
  double v0 = 0.0, v1 = 0.0, v2 = 0.0;
  if (n < (int)g.vol) {
    // planes in the order of the reference's double loop: (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
    const int i = pl < 3 ? 0 : (pl < 5 ? 1 : 2), j = pl < 3 ? pl + 1 : (pl < 5 ? pl - 1 : 3);
    const int e0 = g.ext[0], ne0 = g.ne0;
    int row = n / e0;
    const int lane = n - row * e0;
    const int x0 = lane < ne0 ? 2 * lane : 2 * (lane - ne0) + 1;
    int q = row / g.ext[1];
    const int x1 = row - q * g.ext[1] + g.ghost[1];
    row = q;
    q = row / g.ext[2];
    const int x2 = row - q * g.ext[2] + g.ghost[2];
    const int x3 = q + g.ghost[3];
    const int s1 = (int)g.sstride[1], s2 = (int)g.sstride[2], s3 = (int)g.sstride[3];
    const int sl0 = (x0 & 1) * ne0 + (x0 >> 1);
    const int p = x1 * s1 + x2 * s2 + x3 * s3 + sl0;
    const int x0p = x0 + 1 < e0 ? x0 + 1 : 0;
    const int up0 = (x0p & 1) * ne0 + (x0p >> 1) - sl0;
    const int up1 = x1 + 1 < g.sext[1] ? s1 : -x1 * s1;
    const int up2 = x2 + 1 < g.sext[2] ? s2 : -x2 * s2;
    const int up3 = x3 + 1 < g.sext[3] ? s3 : -x3 * s3;
    const int upi = lq_sel4(i, up0, up1, up2, up3), upj = lq_sel4(j, up0, up1, up2, up3);
    const M3 a = m3_mul_nn(lq_ld36(U, p, i), lq_ld36(U, p + upi, j));
    const M3 b = m3_mul_nn(lq_ld36(U, p, j), lq_ld36(U, p + upj, i));
    const cx t = m3_trace_nd(a, b);
    v0 = t.x;
    v1 = t.y;
    v2 = 1.0 - t.x / CA;
  }
  __shared__ double sm[3][6];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v0 += __shfl_down_sync(0xffffffffu, v0, o);
    v1 += __shfl_down_sync(0xffffffffu, v1, o);
    v2 += __shfl_down_sync(0xffffffffu, v2, o);
  }
  if (lane32 == 0) {
    sm[0][pl] = v0;
    sm[1][pl] = v1;
    sm[2][pl] = v2;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double x = 0.0;
#pragma unroll
    for (int w = 0; w < 6; ++w) x += sm[threadIdx.x][w];
    partial[(lq_i64)blockIdx.x * 3 + threadIdx.x] = x;
  }
}

// ------------------------------------------------------------------------------------------------------------
// AoS <-> chunked-SoA transposition of the links at the C-ABI boundary (lq_links_upload / lq_links_download:
// LatticeStateNew::new, set_link_matrix, link_matrix(), state.rs:779-815), D = 4, ext0 a multiple of 32.
// A row of x0 is one contiguous run on BOTH sides: ext0/32 whole chunks of the device layout (ext0 x 576 B) and ext0
// sites x 4 links x 144 B of the reference's Vec<Matrix3<Complex<f64>>>.  One block per row: a single thread moves the
// row into shared memory with ONE bulk asynchronous copy (TMA, cp.async.bulk + mbarrier transaction count), the block
// permutes it in shared memory (slot -> x0: even sites first; plane-major -> link-major; row-major 3x3 -> nalgebra's
// column-major), and one bulk copy writes it out.  Global memory only ever sees full contiguous rows, where the
// per-thread functors (KLinksToAos / KLinksFromAos) touch the AoS side in 16-byte pieces at a 144-byte stride.
__device__ __forceinline__ unsigned lq_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lq_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lq_smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void lq_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lq_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lq_mbar_wait(unsigned long long* bar, unsigned phase) {
  unsigned ok;
  do {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(lq_smem_u32(bar)), "r"(phase)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void lq_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   lq_smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(lq_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void lq_bulk_s2g(void* gmem_dst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(lq_smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
template <int TO_AOS>
__global__ void __launch_bounds__(128)
    lq_aos4_tma_kernel(LqGeom g, cx* __restrict__ U, double* __restrict__ aos) {
  extern __shared__ __align__(128) unsigned char lq_aos_smem[];
  __shared__ __align__(8) unsigned long long bar;
  const int e0 = g.ext[0], ne0 = g.ne0;
  const unsigned row_bytes = (unsigned)e0 * 576u;
  cx* soa = (cx*)lq_aos_smem;
  cx* lin = (cx*)(lq_aos_smem + row_bytes);
  const int rowi = blockIdx.x;
  int q = rowi / g.ext[1];
  const int x1 = rowi - q * g.ext[1] + g.ghost[1];
  int r2 = q;
  q = r2 / g.ext[2];
  const int x2 = r2 - q * g.ext[2] + g.ghost[2];
  const int x3 = q + g.ghost[3];
  const lq_i64 base = (lq_i64)x1 * g.sstride[1] + (lq_i64)x2 * g.sstride[2] + (lq_i64)x3 * g.sstride[3];
  cx* gsoa = U + (base >> 5) * (36 * 32);
  cx* gaos = (cx*)aos + (lq_i64)rowi * e0 * 36;
  if (threadIdx.x == 0) lq_mbar_init(&bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    lq_mbar_expect_tx(&bar, row_bytes);
    lq_bulk_g2s(TO_AOS ? soa : lin, TO_AOS ? gsoa : gaos, row_bytes, &bar);
  }
  lq_mbar_wait(&bar, 0);
  for (int e = threadIdx.x; e < e0 * 36; e += 128) {
    const int lane = e & 31, pc = e >> 5;   // pc = chunk * 36 + plane
    const int c = pc / 36, plane = pc - c * 36;
    const int slot = c * 32 + lane;
    const int x0 = slot < ne0 ? 2 * slot : 2 * (slot - ne0) + 1;
    const int dir = plane / 9, k = plane - dir * 9;
    const int rr = k / 3, cc = k - rr * 3;
    const int a = (x0 * 4 + dir) * 9 + cc * 3 + rr;  // nalgebra ArrayStorage: column-major
    if (TO_AOS) lin[a] = soa[e];
    else soa[e] = lin[a];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk copy
  __syncthreads();
  if (threadIdx.x == 0) lq_bulk_s2g(TO_AOS ? (void*)gaos : (void*)gsoa, TO_AOS ? lin : soa, row_bytes);
}

// ------------------------------------------------------------------------------------------------------------
// launchers used by lq_capi.cu (D = 4; 32-bit element indices: fields below 2^31 elements, else the generic path)
static inline bool lq_tuned_ok(const LqGeom& g) {
  return g.D == 4 && g.nchunk * 32 * 36 < ((lq_i64)1 << 31) && g.vol < ((lq_i64)1 << 31);
}
static inline cudaError_t lq_tuned_efield_step(cudaStream_t st, const LqGeom& g, const cx* U, cx* E, double coef, double dt,
                                               int nkick) {
  constexpr int BLOCK = 128;
  lq_md4_kernel<BLOCK, 3, 0, 2><<<(unsigned)((g.vol + BLOCK / 4 - 1) / (BLOCK / 4)), BLOCK, 0, st>>>(g, U, nullptr, E, coef, dt,
                                                                                                0.0, 0.0, nkick, nullptr, 0);
  return cudaGetLastError();
}
static inline cudaError_t lq_tuned_efield_link_step(cudaStream_t st, const LqGeom& g, const cx* U, cx* Unew, cx* E,
                                                    double coef, double dt_e, double dt_u, double c_u, int nkick,
                                                    int use_exp = 0) {
  constexpr int BLOCK = 128;
  const unsigned nb = (unsigned)((g.vol + BLOCK / 4 - 1) / (BLOCK / 4));
  if (use_exp)
    lq_md4_kernel<BLOCK, 3, 1, 2 | 256><<<nb, BLOCK, 0, st>>>(g, U, Unew, E, coef, dt_e, dt_u, c_u, nkick, nullptr, 0);
  else
    lq_md4_kernel<BLOCK, 3, 1, 2><<<nb, BLOCK, 0, st>>>(g, U, Unew, E, coef, dt_e, dt_u, c_u, nkick, nullptr, 0);
  return cudaGetLastError();
}
static inline cudaError_t lq_tuned_sweep(cudaStream_t st, const LqGeom& g, cx* U, int kind /*0 hb, 1 or*/, int mu, int parity,
                                         int flags, int or_kind, double coupling, unsigned long long seed,
                                         unsigned long long counter) {
  constexpr int BLOCK = 128;
  const unsigned nb = (unsigned)((g.vol / 2 + BLOCK - 1) / BLOCK);
  if (kind == 0)
    lq_sweep4_kernel<BLOCK, 3, 0><<<nb, BLOCK, 0, st>>>(g, U, mu, parity, flags, or_kind, coupling, seed, counter);
  else
    lq_sweep4_kernel<BLOCK, 3, 1><<<nb, BLOCK, 0, st>>>(g, U, mu, parity, flags, or_kind, coupling, seed, counter);
  return cudaGetLastError();
}
// links AoS <-> SoA through bulk (TMA) copies of whole rows; false: geometry not covered (use the per-thread functors)
static inline bool lq_tuned_aos_ok(const LqGeom& g) { return g.D == 4 && g.ext[0] % 32 == 0 && g.ext[0] <= 128; }
static inline cudaError_t lq_tuned_links_aos(cudaStream_t st, const LqGeom& g, cx* U, double* aos, bool to_aos) {
  const unsigned smem = 2u * (unsigned)g.ext[0] * 576u;
  const unsigned rows = (unsigned)(g.vol / g.ext[0]);
  cudaError_t e;
  if (to_aos) {
    e = cudaFuncSetAttribute(lq_aos4_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    lq_aos4_tma_kernel<1><<<rows, 128, smem, st>>>(g, U, aos);
  } else {
    e = cudaFuncSetAttribute(lq_aos4_tma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    lq_aos4_tma_kernel<0><<<rows, 128, smem, st>>>(g, U, aos);
  }
  return cudaGetLastError();
}
static inline cudaError_t lq_tuned_gauss_field(cudaStream_t st, const LqGeom& g, const cx* U, const cx* E, cx* G,
                                               const LqPush* d_ps) {
  constexpr int BLOCK = 128;
  const unsigned nb = (unsigned)((g.vol + BLOCK - 1) / BLOCK);
  if (d_ps) lq_gfield4_kernel<BLOCK, 3, 1><<<nb, BLOCK, 0, st>>>(g, U, E, G, d_ps);
  else lq_gfield4_kernel<BLOCK, 3, 0><<<nb, BLOCK, 0, st>>>(g, U, E, G, nullptr);
  return cudaGetLastError();
}
static inline cudaError_t lq_tuned_gauss_step(cudaStream_t st, const LqGeom& g, const cx* U, const cx* G, const cx* Ein,
                                              cx* Eout) {
  constexpr int BLOCK = 128;
  // 128 registers / 4 blocks per SM: 0.206 ms at 32^4; 166 / 3: 0.222; 96 / 5 (spills): 0.236; 80 / 6: 0.338
  lq_gstep4_kernel<BLOCK, 4><<<(unsigned)((g.vol + BLOCK / 4 - 1) / (BLOCK / 4)), BLOCK, 0, st>>>(g, U, G, Ein, Eout);
  return cudaGetLastError();
}
static inline lq_i64 lq_tuned_metropolis_blocks(const LqGeom& g) { return (g.vol / 2 + 127) / 128; }
static inline cudaError_t lq_tuned_metropolis(cudaStream_t st, const LqGeom& g, cx* U, int mu, int parity, int flags,
                                              int n_update, double beta, double CA, double spread, unsigned long long seed,
                                              unsigned long long counter, double* partial) {
  lq_metro4_kernel<128, 3><<<(unsigned)lq_tuned_metropolis_blocks(g), 128, 0, st>>>(g, U, mu, parity, flags, n_update, beta,
                                                                                  CA, spread, seed, counter, partial);
  return cudaGetLastError();
}
// per-block partial sums of the plaquette terms: ceil(vol / 32) blocks x 3 doubles
static inline lq_i64 lq_tuned_plaquette_blocks(const LqGeom& g) { return (g.vol + 31) / 32; }
static inline cudaError_t lq_tuned_plaquette(cudaStream_t st, const LqGeom& g, const cx* U, double CA, double* partial) {
  // register budgets of 168 / 113 / 85 / 68 per thread (2..5 blocks per SM) all run in 0.31-0.39 ms at 32^4; ncu: no unit
  // above half, 43 % of the instructions are site decode + block reduction (profiles/r01zi_plaq4_ncu.txt)
  lq_plaq4_kernel<2><<<(unsigned)lq_tuned_plaquette_blocks(g), 192, 0, st>>>(g, U, CA, partial);
  return cudaGetLastError();
}
// fused force + E kick + link step + halo push of the new boundary links into the neighbours' ghost layers
static inline cudaError_t lq_tuned_efield_link_step_push(cudaStream_t st, const LqGeom& g, const cx* U, cx* Unew, cx* E,
                                                         double coef, double dt_e, double dt_u, double c_u, int nkick,
                                                         const LqPush* d_ps, int use_exp = 0) {
  constexpr int BLOCK = 128;
  const lq_i64 slice = g.vol / g.ext[3];
  const int bps = (g.ghost[3] && slice % (BLOCK / 4) == 0) ? (int)(slice / (BLOCK / 4)) : 0;
  const unsigned nb = (unsigned)((g.vol + BLOCK / 4 - 1) / (BLOCK / 4));
  if (use_exp)
    lq_md4_kernel<BLOCK, 3, 1, 2 | 256, 1><<<nb, BLOCK, 0, st>>>(g, U, Unew, E, coef, dt_e, dt_u, c_u, nkick, d_ps, bps);
  else
    lq_md4_kernel<BLOCK, 3, 1, 2, 1><<<nb, BLOCK, 0, st>>>(g, U, Unew, E, coef, dt_e, dt_u, c_u, nkick, d_ps, bps);
  return cudaGetLastError();
}

#endif  // !LQ_HOST_EMU
