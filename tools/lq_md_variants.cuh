// tools/lq_md_variants.cuh -- kernel variants of the MD hot loop that were measured and NOT adopted, kept for
// tools/kbench.cu only (development tool; nothing here is compiled into liblqcd_b200.so).
// Every variant is bit-compared with the generic functor KEfieldLinkStep by kbench before it is timed; the measured
// records are under profiles/ (r01b-f, i, q, u, w and r02*).
#pragma once
#include "lq_tuned.cuh"

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
__device__ __forceinline__ void lq_md4x_body(const LqGeom& g, const cx* __restrict__ U, cx* __restrict__ Unew,
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
    lq_md4x_kernel(LqGeom g, const cx* __restrict__ U, cx* __restrict__ Unew, cx* __restrict__ E, double coef, double dt_e,
                  double dt_u, double c_u, int nkick, const LqPush* __restrict__ ps, int bps) {
  // ps: device-resident peer table, read by the threads of boundary slices only; bps: blocks per t-slice (0: keep
  // the natural block order)
  if (FLAGS & 16) {
    constexpr int SITES = BLOCK / 4;
    const int nblk = ((int)g.vol + SITES - 1) / SITES;
    const int per = (nblk + (int)gridDim.x - 1) / (int)gridDim.x;
    const int b0 = blockIdx.x * per, b1 = min(b0 + per, nblk);
#pragma unroll 1
    for (int b = b0; b < b1; ++b) lq_md4x_body<BLOCK, FUSED, FLAGS, 0>(g, U, Unew, E, coef, dt_e, dt_u, c_u, nkick, ps, b);
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
  lq_md4x_body<BLOCK, FUSED, FLAGS, PUSH>(g, U, Unew, E, coef, dt_e, dt_u, c_u, nkick, ps, blk);
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
// V6: product-level software pipeline with a SMALL live set.  The staple sum is a chain of 12 products
//   t = a b^+, acc += t c^+   (up)      t = a b, acc += t^+ c   (down)
// and the operands of the next product are requested while the current one is computed: live = acc + t + a + b + c
// = 90 registers (V5 pipelined whole staples: 126 + 36).  The point is to fit 4 (128 registers) or 5 (96) warps per
// scheduler with every load one product (216 DFMAs) ahead of its first use.
template <int BLOCK, int MINB, int FUSED>
__global__ void __launch_bounds__(BLOCK, MINB)
    lq_md6_kernel(LqGeom g, const cx* __restrict__ U, cx* __restrict__ Unew, cx* __restrict__ E, double coef, double dt_e,
                  double dt_u, double c_u, int nkick) {
  constexpr int SITES = BLOCK / 4;
  const int mu = threadIdx.x / SITES;
  const int n = blockIdx.x * SITES + (threadIdx.x - mu * SITES);
  if (n >= (int)g.vol) return;
  const LqSite4 s = lq_site4(g, n);
  const int p = s.p;
  const int pm = p + lq_sel4(mu, s.up[0], s.up[1], s.up[2], s.up[3]);
  const int ee = ((p >> 5) * 16 + mu * 4) * 32 + (p & 31);
  M3 acc = m3_zero();
  int nu = (mu + 1) & 3;
  int upn = lq_sel4(nu, s.up[0], s.up[1], s.up[2], s.up[3]), dnn = lq_sel4(nu, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
  M3 a = lq_ld36(U, pm, nu);
  M3 b = lq_ld36(U, p + upn, mu);
  cx ev[4];
#pragma unroll 1
  for (int j = 1; j < 4; ++j) {
    M3 c = lq_ld36(U, p, nu);
    M3 t = m3_mul_nd(a, b);               // up: U_nu(x+mu) U_mu^+(x+nu)
    a = lq_ld36(U, p + dnn, mu);
    b = lq_ld36(U, pm + dnn, nu);
    m3_fma_nd(acc, t, c);                 //     ... U_nu^+(x)
    c = lq_ld36(U, p + dnn, nu);
    t = m3_mul_nn(a, b);                  // down: U_mu(x-nu) U_nu(x+mu-nu)
    if (j < 3) {
      nu = (mu + j + 1) & 3;
      upn = lq_sel4(nu, s.up[0], s.up[1], s.up[2], s.up[3]);
      dnn = lq_sel4(nu, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
      a = lq_ld36(U, pm, nu);
      b = lq_ld36(U, p + upn, mu);
    } else {
      a = lq_ld36(U, p, mu);              // the link itself, for U * A
#pragma unroll
      for (int k = 0; k < 4; ++k) ev[k] = __ldcs(E + ee + k * 32);
    }
    m3_fma_dn(acc, t, c);                 //     (...)^+ U_nu(x-nu)
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
// V7: V6 written as straight-line code (the three nu pairs unrolled by hand, operands requested one product ahead of
// their first use).  ptxas keeps the whole schedule in 128 registers without spills (16 warps per SM) and interleaves
// the loads of the next product with the DFMAs of the current one.
// straight-line product-level pipeline (fully unrolled), loads one product ahead
template <int BLOCK, int MINB, int FUSED>
__global__ void __launch_bounds__(BLOCK, MINB)
    lq_md7_kernel(LqGeom g, const cx* __restrict__ U, cx* __restrict__ Unew, cx* __restrict__ E, double coef, double dt_e,
                  double dt_u, double c_u, int nkick) {
  constexpr int SITES = BLOCK / 4;
  const int mu = threadIdx.x / SITES;
  const int n = blockIdx.x * SITES + (threadIdx.x - mu * SITES);
  if (n >= (int)g.vol) return;
  const LqSite4 s = lq_site4(g, n);
  const int p = s.p;
  const int pm = p + lq_sel4(mu, s.up[0], s.up[1], s.up[2], s.up[3]);
  const int ee = ((p >> 5) * 16 + mu * 4) * 32 + (p & 31);
  M3 acc = m3_zero();
  const int n1 = (mu + 1) & 3, n2 = (mu + 2) & 3, n3 = (mu + 3) & 3;
  const int u1 = lq_sel4(n1, s.up[0], s.up[1], s.up[2], s.up[3]), d1 = lq_sel4(n1, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
  const int u2 = lq_sel4(n2, s.up[0], s.up[1], s.up[2], s.up[3]), d2 = lq_sel4(n2, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
  const int u3 = lq_sel4(n3, s.up[0], s.up[1], s.up[2], s.up[3]), d3 = lq_sel4(n3, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
  M3 a = lq_ld36(U, pm, n1);
  M3 b = lq_ld36(U, p + u1, mu);
  M3 c, t;
#define STAGE_UP(NU, DN)                   \
  c = lq_ld36(U, p, NU);                   \
  t = m3_mul_nd(a, b);                     \
  a = lq_ld36(U, p + DN, mu);              \
  b = lq_ld36(U, pm + DN, NU);             \
  m3_fma_nd(acc, t, c);
#define STAGE_DN(NU, DN, NEXTA, NEXTB)     \
  c = lq_ld36(U, p + DN, NU);              \
  t = m3_mul_nn(a, b);                     \
  a = NEXTA;                               \
  b = NEXTB;                               \
  m3_fma_dn(acc, t, c);
  STAGE_UP(n1, d1)
  STAGE_DN(n1, d1, lq_ld36(U, pm, n2), lq_ld36(U, p + u2, mu))
  STAGE_UP(n2, d2)
  STAGE_DN(n2, d2, lq_ld36(U, pm, n3), lq_ld36(U, p + u3, mu))
  STAGE_UP(n3, d3)
  cx ev[4];
  c = lq_ld36(U, p + d3, n3);
  t = m3_mul_nn(a, b);
  a = lq_ld36(U, p, mu);
#pragma unroll
  for (int k = 0; k < 4; ++k) ev[k] = __ldcs(E + ee + k * 32);
  m3_fma_dn(acc, t, c);
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
// V8: row split -- three threads per link, thread r carries row r of the running product:
//   t_r = a_r b^+,  acc_r += t_r c^+   (up)       t_r = a_r b,  s_r = t_r ... (down needs (a b)^+ c: column access)
// The down staple (a b)^+ c = b^+ a^+ c is evaluated as row r of b^+ (= conj of column r of b) times a^+, times c, so
// every stage is "row vector x matrix": 6 + 18 live operand registers instead of 36.  Each thread still needs ALL of
// the two right-hand matrices of a stage, so the register-level operand traffic is 7/3 of the one-thread-per-link
// mapping (the three row threads sit in three different warps and read the same 512-byte lines through L1).
// Block = 32 sites x 4 mu x 3 rows = 384 threads; the three rows of the staple sum meet in shared memory and the
// row-0 thread finishes the link (U A, trace, E kick, link step) exactly as the other variants do.
__device__ __forceinline__ void lq_row_nd(cx r[3], const cx v[3], const M3& m) {  // r = v * m^+
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    cx s = cmk(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < 3; ++k) cfma_c(s, v[k], m.e[3 * j + k]);
    r[j] = s;
  }
}
__device__ __forceinline__ void lq_row_nn(cx r[3], const cx v[3], const M3& m) {  // r = v * m
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    cx s = cmk(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < 3; ++k) cfma(s, v[k], m.e[3 * k + j]);
    r[j] = s;
  }
}
template <int MINB, int FUSED>
__global__ void __launch_bounds__(384, MINB)
    lq_md8_rowsplit_kernel(LqGeom g, const cx* __restrict__ U, cx* __restrict__ Unew, cx* __restrict__ E, double coef,
                           double dt_e, double dt_u, double c_u, int nkick) {
  __shared__ cx sm[4][2][3][32];  // rows 1 and 2 of the staple sum, per direction
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;  // 0..11
  const int mu = w / 3, r = w - 3 * mu;
  const int n = blockIdx.x * 32 + lane;
  const bool live = n < (int)g.vol;
  cx acc[3] = {cmk(0, 0), cmk(0, 0), cmk(0, 0)};
  LqSite4 s;
  int p = 0;
  if (live) {
    s = lq_site4(g, n);
    p = s.p;
    const int pm = p + lq_sel4(mu, s.up[0], s.up[1], s.up[2], s.up[3]);
#pragma unroll 1
    for (int j = 1; j < 4; ++j) {
      const int nu = (mu + j) & 3;
      const int upn = lq_sel4(nu, s.up[0], s.up[1], s.up[2], s.up[3]), dnn = lq_sel4(nu, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
      {  // up: row r of U_nu(x+mu), times U_mu^+(x+nu), times U_nu^+(x)
        const cx* ab = U + ((pm >> 5) * 36 + nu * 9 + 3 * r) * 32 + (pm & 31);
        cx v[3] = {__ldg(ab), __ldg(ab + 32), __ldg(ab + 64)}, t[3], q[3];
        lq_row_nd(t, v, lq_ld36(U, p + upn, mu));
        lq_row_nd(q, t, lq_ld36(U, p, nu));
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[k] = cadd(acc[k], q[k]);
      }
      {  // down: row r of U_nu^+(x+mu-nu) = conj of column r of U_nu(x+mu-nu), times U_mu^+(x-nu), times U_nu(x-nu)
        const int pb = pm + dnn;
        const cx* bb = U + ((pb >> 5) * 36 + nu * 9 + r) * 32 + (pb & 31);
        cx v[3] = {cconj(__ldg(bb)), cconj(__ldg(bb + 96)), cconj(__ldg(bb + 192))}, t[3], q[3];
        lq_row_nd(t, v, lq_ld36(U, p + dnn, mu));
        lq_row_nn(q, t, lq_ld36(U, p + dnn, nu));
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[k] = cadd(acc[k], q[k]);
      }
    }
    if (r > 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) sm[mu][r - 1][k][lane] = acc[k];
    }
  }
  __syncthreads();
  if (!live || r != 0) return;
  M3 a;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    a.e[k] = acc[k];
    a.e[3 + k] = sm[mu][0][k][lane];
    a.e[6 + k] = sm[mu][1][k][lane];
  }
  const M3 u = lq_ld36(U, p, mu);
  M3 wm = m3_mul_nn(u, a);
  cx tr[8];
  lq_trace_gen(wm, tr);
  const int ee = ((p >> 5) * 16 + mu * 4) * 32 + (p & 31);
  A8 e;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const cx v = __ldcs(E + ee + k * 32);
    e.e[2 * k] = v.x;
    e.e[2 * k + 1] = v.y;
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

#undef STAGE_UP
#undef STAGE_DN
// ------------------------------------------------------------------------------------------------------------
// V9: V7 with every 3x3 product fenced into its own basic block (a one-trip loop whose bound is a kernel argument):
// ptxas cannot interleave the DFMAs of two products or slide loads into the middle of one, so (i) the operands of the
// next product are requested exactly one product ahead, at the block boundary, and (ii) inside an accumulating product
// the operand-stationary FMA order of lq_common.cuh survives (runs of six DFMAs sharing a multiplicand: .reuse on 83 %
// of them; the products that start from zero still get reordered, ~50 %).
// straight-line product pipeline, every product fenced into its own basic block (opaque one-trip loop)
#define FENCE_BEGIN _Pragma("unroll 1") for (int f_ = 0; f_ < one; ++f_) {
#define FENCE_END }
template <int BLOCK, int MINB, int FUSED>
__global__ void __launch_bounds__(BLOCK, MINB)
    lq_md9_kernel(LqGeom g, const cx* __restrict__ U, cx* __restrict__ Unew, cx* __restrict__ E, double coef, double dt_e,
                  double dt_u, double c_u, int nkick, int one) {
  constexpr int SITES = BLOCK / 4;
  const int mu = threadIdx.x / SITES;
  const int n = blockIdx.x * SITES + (threadIdx.x - mu * SITES);
  if (n >= (int)g.vol) return;
  const LqSite4 s = lq_site4(g, n);
  const int p = s.p;
  const int pm = p + lq_sel4(mu, s.up[0], s.up[1], s.up[2], s.up[3]);
  const int ee = ((p >> 5) * 16 + mu * 4) * 32 + (p & 31);
  M3 acc = m3_zero();
  const int n1 = (mu + 1) & 3, n2 = (mu + 2) & 3, n3 = (mu + 3) & 3;
  const int u1 = lq_sel4(n1, s.up[0], s.up[1], s.up[2], s.up[3]), d1 = lq_sel4(n1, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
  const int u2 = lq_sel4(n2, s.up[0], s.up[1], s.up[2], s.up[3]), d2 = lq_sel4(n2, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
  const int u3 = lq_sel4(n3, s.up[0], s.up[1], s.up[2], s.up[3]), d3 = lq_sel4(n3, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
  M3 a = lq_ld36(U, pm, n1);
  M3 b = lq_ld36(U, p + u1, mu);
  M3 c, t;
#define STAGE_UP(NU, DN)                   \
  c = lq_ld36(U, p, NU);                   \
  FENCE_BEGIN t = m3_mul_nd(a, b); FENCE_END \
  a = lq_ld36(U, p + DN, mu);              \
  b = lq_ld36(U, pm + DN, NU);             \
  FENCE_BEGIN m3_fma_nd(acc, t, c); FENCE_END
#define STAGE_DN(NU, DN, NEXTA, NEXTB)     \
  c = lq_ld36(U, p + DN, NU);              \
  FENCE_BEGIN t = m3_mul_nn(a, b); FENCE_END \
  a = NEXTA;                               \
  b = NEXTB;                               \
  FENCE_BEGIN m3_fma_dn(acc, t, c); FENCE_END
  STAGE_UP(n1, d1)
  STAGE_DN(n1, d1, lq_ld36(U, pm, n2), lq_ld36(U, p + u2, mu))
  STAGE_UP(n2, d2)
  STAGE_DN(n2, d2, lq_ld36(U, pm, n3), lq_ld36(U, p + u3, mu))
  STAGE_UP(n3, d3)
  cx ev[4];
  c = lq_ld36(U, p + d3, n3);
  FENCE_BEGIN t = m3_mul_nn(a, b); FENCE_END
  a = lq_ld36(U, p, mu);
#pragma unroll
  for (int k = 0; k < 4; ++k) ev[k] = __ldcs(E + ee + k * 32);
  FENCE_BEGIN m3_fma_dn(acc, t, c); FENCE_END
  const M3 u = a;
  M3 w;
  FENCE_BEGIN w = m3_mul_nn(u, acc); FENCE_END
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
#undef FENCE_BEGIN
#undef FENCE_END
#undef STAGE_UP
#undef STAGE_DN

// ------------------------------------------------------------------------------------------------------------
// V10: V7 with the (a, b) operand pairs requested TWO staples ahead (two register sets that alternate) and c one
// product ahead: live = acc + t + 2 (a, b) + c = 126 registers, for the 168-register / 12-warp budget.
// NUORD: 0 = nu = mu+1, mu+2, mu+3 (mod 4); 1 = ascending with the own direction skipped.
template <int BLOCK, int MINB, int FUSED, int NUORD = 0>
__global__ void __launch_bounds__(BLOCK, MINB)
    lq_md10_kernel(LqGeom g, const cx* __restrict__ U, cx* __restrict__ Unew, cx* __restrict__ E, double coef, double dt_e,
                   double dt_u, double c_u, int nkick) {
  constexpr int SITES = BLOCK / 4;
  const int mu = threadIdx.x / SITES;
  const int n = blockIdx.x * SITES + (threadIdx.x - mu * SITES);
  if (n >= (int)g.vol) return;
  const LqSite4 s = lq_site4(g, n);
  const int p = s.p;
  const int pm = p + lq_sel4(mu, s.up[0], s.up[1], s.up[2], s.up[3]);
  const int ee = ((p >> 5) * 16 + mu * 4) * 32 + (p & 31);
  M3 acc = m3_zero();
  const int n1 = NUORD ? (mu == 0 ? 1 : 0) : (mu + 1) & 3;
  const int n2 = NUORD ? (mu <= 1 ? 2 : 1) : (mu + 2) & 3;
  const int n3 = NUORD ? (mu <= 2 ? 3 : 2) : (mu + 3) & 3;
  const int u1 = lq_sel4(n1, s.up[0], s.up[1], s.up[2], s.up[3]), d1 = lq_sel4(n1, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
  const int u2 = lq_sel4(n2, s.up[0], s.up[1], s.up[2], s.up[3]), d2 = lq_sel4(n2, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
  const int u3 = lq_sel4(n3, s.up[0], s.up[1], s.up[2], s.up[3]), d3 = lq_sel4(n3, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
  // staple operands:  up(nu):  a = U_nu(x+mu), b = U_mu(x+nu), c = U_nu(x);   t = a b^+, acc += t c^+
  //                   dn(nu):  a = U_mu(x-nu), b = U_nu(x+mu-nu), c = U_nu(x-nu);  t = a b, acc += t^+ c
  M3 a0 = lq_ld36(U, pm, n1), b0 = lq_ld36(U, p + u1, mu);
  M3 a1 = lq_ld36(U, p + d1, mu), b1 = lq_ld36(U, pm + d1, n1);
  M3 c = lq_ld36(U, p, n1), t;
  t = m3_mul_nd(a0, b0);                                            // up 1
  a0 = lq_ld36(U, pm, n2);
  b0 = lq_ld36(U, p + u2, mu);
  m3_fma_nd(acc, t, c);
  c = lq_ld36(U, p + d1, n1);
  t = m3_mul_nn(a1, b1);                                            // down 1
  a1 = lq_ld36(U, p + d2, mu);
  b1 = lq_ld36(U, pm + d2, n2);
  m3_fma_dn(acc, t, c);
  c = lq_ld36(U, p, n2);
  t = m3_mul_nd(a0, b0);                                            // up 2
  a0 = lq_ld36(U, pm, n3);
  b0 = lq_ld36(U, p + u3, mu);
  m3_fma_nd(acc, t, c);
  c = lq_ld36(U, p + d2, n2);
  t = m3_mul_nn(a1, b1);                                            // down 2
  a1 = lq_ld36(U, p + d3, mu);
  b1 = lq_ld36(U, pm + d3, n3);
  m3_fma_dn(acc, t, c);
  c = lq_ld36(U, p, n3);
  t = m3_mul_nd(a0, b0);                                            // up 3
  a0 = lq_ld36(U, p, mu);                                           // the link itself, for U * A
  cx ev[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) ev[k] = __ldcs(E + ee + k * 32);
  m3_fma_nd(acc, t, c);
  c = lq_ld36(U, p + d3, n3);
  t = m3_mul_nn(a1, b1);                                            // down 3
  m3_fma_dn(acc, t, c);
  const M3 u = a0;
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
// Sweep sub-step with the ROLLED nu loop of round 1 (A/B partner of lq_sweep4_kernel, whose staple sum is the
// straight-line pipeline lq_staples4)
template <int BLOCK, int MINB, int KIND>
__global__ void __launch_bounds__(BLOCK, MINB)
    lq_sweep4_rolled_kernel(LqGeom g, cx* __restrict__ U, int mu, int parity, int flags, int or_kind, double coupling,
                            unsigned long long seed, unsigned long long counter) {
  const int n = blockIdx.x * BLOCK + threadIdx.x;
  if (n >= (int)(g.vol >> 1)) return;
  lq_i64 gi;
  const LqSite4 s = lq_site4_eo(g, n, parity, gi);
  const int p = s.p;
  const int pm = p + lq_sel4(mu, s.up[0], s.up[1], s.up[2], s.up[3]);
  M3 acc = m3_zero();
#pragma unroll 1
  for (int j = 0; j < 3; ++j) {
    const int nu = j + (j >= mu ? 1 : 0);
    const int upn = lq_sel4(nu, s.up[0], s.up[1], s.up[2], s.up[3]), dnn = lq_sel4(nu, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
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
    LqStream rng(seed, counter, (uint64_t)(gi * 4 + mu));
    r = lq_heat_bath_link(u, acc, coupling, rng, flags);
  } else {
    r = lq_overrelax_link(u, acc, or_kind);
  }
#pragma unroll
  for (int kk = 0; kk < 9; ++kk) own[kk * 32] = r.e[kk];
}
// staple phase ALONE (no rule; writes the staple sum over the own link: wrong on purpose, timing only): what a sweep
// sub-step costs before the single-link rule
template <int BLOCK, int MINB, int PIPE>
__global__ void __launch_bounds__(BLOCK, MINB)
    lq_sweep4_staples_only_kernel(LqGeom g, cx* __restrict__ U, cx* __restrict__ out, int mu, int parity) {
  const int n = blockIdx.x * BLOCK + threadIdx.x;
  if (n >= (int)(g.vol >> 1)) return;
  lq_i64 gi;
  const LqSite4 s = lq_site4_eo(g, n, parity, gi);
  cx* own = U + ((s.p >> 5) * 36 + mu * 9) * 32 + (s.p & 31);
  M3 acc, u;
  lq_staples4(U, own, s, mu, acc, u);
  cx* o = out + ((s.p >> 5) * 36 + mu * 9) * 32 + (s.p & 31);
#pragma unroll
  for (int kk = 0; kk < 9; ++kk) o[kk * 32] = cadd(acc.e[kk], u.e[kk]);
}

// ------------------------------------------------------------------------------------------------------------
// The same sub-step, WARP SPECIALISED: staple warps and rule warps of one persistent block run concurrently.
// lq_sweep4_kernel serialises two very different phases in every thread -- the staple sum (HBM-bound, 168 registers,
// 0.97 ms of a 1.47 ms heat-bath sweep at 32^4) and the single-link rule (scalar dependency chains: Philox rounds,
// log, cospi, square roots; needs the link and a few temporaries) -- at the occupancy of the hungrier one, 3 warps per
// scheduler.  Here the block is split: NPW staple warps compute the staple sums of 32-link tasks and hand (A, U) to
// NCW rule warps through a ring of shared-memory slots (one mbarrier pair per slot: full / free; slot s % RING is
// always written by staple warp s % NPW and read by rule warp s % NCW, so the schedule is static); a rule warp draws,
// updates and stores the links of one task while the staple warps stream the next ones.  The register file is
// re-partitioned at the start (setmaxnreg: PREG for the staple warps, CREG for the rule warps; setmaxnreg.inc only
// draws from what the block's own warps released: NPW (PREG - R0) <= NCW (R0 - CREG), R0 = the launch allocation).
// MEASURED AND REJECTED (profiles/r02n_kbench2_ws*.txt, 32^4, ms per heat-bath sweep; serial product kernel 1.476):
//   block-wide named barriers, tiles of 256 links, 8 + 8 warps: 1.64 (152/104 registers) ... 1.77 (168/88);
//   this mbarrier ring: 8 + 8 warps 2.19 - 2.41, 12 + 8 warps 2.60 - 2.82, 12 + 4 warps 2.66 - 3.16.
// Why: the staple phase is limited by the bytes its warps keep in flight, and those live in registers -- the register
// file is the prefetch buffer (staple phase alone: 0.97 / 1.04 / 1.12 ms at 12 x 168 / 8 x 244 / 16 x 128 registers,
// i.e. whatever the split of the same 64 K registers).  Handing a third of the file to rule warps costs the staple
// side more than the overlap returns; in the serial kernel the warps drift out of phase anyway, so staple and rule
// phases of different warps already overlap.
// Same arithmetic, order and random streams as lq_sweep4_kernel: bit-identical links.  Heat bath (KIND 0) and the
// SU(2)-sub-group over-relaxation (KIND 1, or_kind 2) read the staple sum from the slot on demand (two columns per
// sub-group); the SVD over-relaxations load it whole.
__device__ __forceinline__ void lq_mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
#define LQ_WS_SLOT (18 * 32) /* cx per slot: (A, U) of 32 links, entry k of lane l at [k * 32 + l] */
template <int NPW, int NCW>
struct LqWsCfg {
  static constexpr int RING = (NPW == 12 && NCW == 8) ? 24 : (NPW == 8 && NCW == 8) ? 24 : (NPW == 12 && NCW == 4) ? 24 : 0;
  static constexpr int THREADS = 32 * (NPW + NCW);
  static constexpr int SMEM = RING * LQ_WS_SLOT * 16 + 2 * RING * 8;
};
// task of sequence number sq inside a block: groups of four adjacent tasks (eight x0 rows: their staples share
// neighbour rows through L1) are dealt to the blocks round robin, so that all blocks work on the same region of the
// lattice at the same time (L2 locality) and the tail is a fraction of one group per block
template <int NPW>
__device__ __forceinline__ int lq_ws_task(int sq) {
  const int w = sq % NPW;
  const int gs = (sq / NPW) * (NPW / 4) + (w >> 2);
  return ((int)blockIdx.x + gs * (int)gridDim.x) * 4 + (w & 3);
}
template <int KIND, int NPW, int NCW, int PREG, int CREG>
__global__ void __launch_bounds__(32 * (NPW + NCW), 1)
    lq_sweep4ws_kernel(LqGeom g, cx* __restrict__ U, int mu, int parity, int flags, int or_kind, double coupling,
                       unsigned long long seed, unsigned long long counter, int ntasks) {
  constexpr int RING = LqWsCfg<NPW, NCW>::RING;
  static_assert(RING > 0 && RING % NPW == 0 && RING % NCW == 0, "ring must pair staple and rule warps statically");
  extern __shared__ __align__(16) unsigned char lq_ws_smem[];
  cx* S = (cx*)lq_ws_smem;
  unsigned long long* full = (unsigned long long*)(lq_ws_smem + RING * LQ_WS_SLOT * 16);
  unsigned long long* empty = full + RING;
  const int half = (int)(g.vol >> 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < RING) {
    lq_mbar_init(&full[threadIdx.x], 32);
    lq_mbar_init(&empty[threadIdx.x], 32);
  }
  __syncthreads();
  if (warp < NPW) {
    if (PREG) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PREG));
    for (int sq = warp;; sq += NPW) {  // sq: sequence number of the task inside this block
      const int task = lq_ws_task<NPW>(sq);
      if (task >= ntasks) break;
      const int slot = sq % RING, use = sq / RING;
      if (use > 0) lq_mbar_wait(&empty[slot], (unsigned)(use - 1) & 1u);
      const int n = task * 32 + lane;
      if (n < half) {
        lq_i64 gi;
        const LqSite4 s = lq_site4_eo(g, n, parity, gi);
        const cx* own = U + ((s.p >> 5) * 36 + mu * 9) * 32 + (s.p & 31);
        M3 acc, u;
        lq_staples4(U, own, s, mu, acc, u);
        cx* d = S + slot * LQ_WS_SLOT + lane;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          d[k * 32] = acc.e[k];
          d[(9 + k) * 32] = u.e[k];
        }
      }
      lq_mbar_arrive(&full[slot]);
    }
  } else {
    if (CREG) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CREG));
    for (int sq = warp - NPW;; sq += NCW) {
      const int task = lq_ws_task<NPW>(sq);
      if (task >= ntasks) break;
      const int slot = sq % RING, use = sq / RING;
      lq_mbar_wait(&full[slot], (unsigned)use & 1u);
      const int n = task * 32 + lane;
      if (n < half) {
        lq_i64 gi;
        const LqSite4 s = lq_site4_eo(g, n, parity, gi);
        cx* own = U + ((s.p >> 5) * 36 + mu * 9) * 32 + (s.p & 31);
        const cx* a = S + slot * LQ_WS_SLOT + lane;
        M3 u, r;
#pragma unroll
        for (int k = 0; k < 9; ++k) u.e[k] = a[(9 + k) * 32];
        if (KIND == 0) {
          LqStream rng(seed, counter, (uint64_t)(gi * 4 + mu));
          r = lq_subgroup_update_acc(u, LqStapleShared<32>{a}, LqHeatBathRule{coupling, &rng, flags});
        } else if (or_kind == 2) {
          r = lq_subgroup_update_acc(u, LqStapleShared<32>{a}, LqOverrelaxSu2Rule{});
        } else {
          M3 acc;
#pragma unroll
          for (int k = 0; k < 9; ++k) acc.e[k] = a[k * 32];
          r = lq_overrelax_link(u, acc, or_kind);
        }
#pragma unroll
        for (int kk = 0; kk < 9; ++kk) own[kk * 32] = r.e[kk];
      }
      lq_mbar_arrive(&empty[slot]);
    }
  }
}


// ------------------------------------------------------------------------------------------------------------
// V11: the TMA-STAGED form north_star names -- neighbour links staged in shared memory by bulk asynchronous copies.
// One persistent block per SM walks the x0 rows of the lattice (ext0 = 32: a row is one 32-slot chunk, so every
// (row, direction) operand of the stencil is ONE contiguous 4608-byte block of the chunked-SoA layout; x0 shifts are
// lane permutations inside a block).  A producer thread issues cp.async.bulk copies armed on mbarrier transaction
// counts; eight consumer warps (direction mu x {up staples, down staples}) read their operands with LDS.128:
//   * per row: the 4 own blocks (double buffered), then three ROUNDS, one per perfect matching of the directions --
//     (0,1)(2,3), (0,2)(1,3), (0,3)(1,2).  In a round every direction's warps work on the plane shared with their
//     partner, so both warps of a plane use the same neighbour blocks: 11 blocks per round (tools/gen_md11_tables.py),
//     37 per row instead of the 76 operand fetches of the one-thread-per-link kernels (31 distinct blocks exist);
//   * three stage buffers (one per round) + full / empty mbarriers: the producer runs up to a whole row ahead;
//   * the down-staple warp hands its partial sum to the up-staple warp of the same direction through shared memory,
//     which finishes the link (U A, trace, E kick, link step) as every other variant does.
// Differences from the generic functor: the staples are summed in round order and as (up sum) + (down sum): errE / errU
// are rounding-level, not zero.  Needs ext0 = 32 and no ghost layers (kbench lattices).
#include "lq_md11_tables.inc"
#define LQ11_BLK 288 /* cx per (row, direction) block: 9 planes x 32 slots */
#define LQ11_SMEM ((2 * 4 + 3 * 11 + 4) * LQ11_BLK * 16 + 32 * 8)
__device__ __forceinline__ unsigned lq_mbar_try(unsigned long long* bar, unsigned phase) {
  unsigned ok;
  asm volatile(
      "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
      : "=r"(ok)
      : "r"(lq_smem_u32(bar)), "r"(phase)
      : "memory");
  return ok;
}
__device__ __forceinline__ void lq_mbar_wait_wd(unsigned long long* bar, unsigned phase) {
  if (lq_mbar_try(bar, phase)) return;
  const long long t0 = clock64();
  while (!lq_mbar_try(bar, phase))
    if (clock64() - t0 > (1ll << 31)) __trap();  // ~1 s: a lost arrival must not hang the box
}
__device__ __forceinline__ M3 lq11_lds(const cx* blk, int ln) {
  M3 r;
#pragma unroll
  for (int k = 0; k < 9; ++k) r.e[k] = blk[k * 32 + ln];
  return r;
}
// (9 warps put 3 on one SM sub-partition: 3 x 32 x registers <= 16 K caps the kernel at 168 registers per thread)
template <int FUSED>
__global__ void __launch_bounds__(288, 1)
    lq_md11_tma_kernel(LqGeom g, const cx* __restrict__ U, cx* __restrict__ Unew, cx* __restrict__ E, double coef,
                       double dt_e, double dt_u, double c_u, int nkick, int ntiles) {
  extern __shared__ __align__(128) unsigned char lq11_smem[];
  cx* own = (cx*)lq11_smem;                 // [2][4][288]
  cx* stage = own + 2 * 4 * LQ11_BLK;       // [3][11][288]
  cx* part = stage + 3 * 11 * LQ11_BLK;     // [4][288]
  unsigned long long* bars = (unsigned long long*)(part + 4 * LQ11_BLK);
  unsigned long long *ownfull = bars, *ownempty = bars + 2, *full = bars + 4, *empty = bars + 7, *partfull = bars + 10,
                     *partempty = bars + 14;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      lq_mbar_init(&ownfull[i], 1);
      lq_mbar_init(&ownempty[i], 8);
    }
    for (int i = 0; i < 3; ++i) {
      lq_mbar_init(&full[i], 1);
      lq_mbar_init(&empty[i], 8);
    }
    for (int i = 0; i < 4; ++i) {
      lq_mbar_init(&partfull[i], 1);
      lq_mbar_init(&partempty[i], 1);
    }
  }
  __syncthreads();
  const int e1 = g.ext[1], e2 = g.ext[2], e3 = g.ext[3];
  if (warp == 8) {
    // ---------------- producer warp: lane j < 11 issues stage block j of every round (its three row offsets and
    // directions are loop invariants), lane 0 also arms the barriers and copies the 4 contiguous own blocks.
    // (One thread issuing all 37 copies of a row, index arithmetic included, took ~3 us per row: 1.11 ms per launch.)
    int o1[3], o2[3], o3[3], dd[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int j = lane < 11 ? lane : 0;
      o1[r] = LQ11_OFF[r][j][0];
      o2[r] = LQ11_OFF[r][j][1];
      o3[r] = LQ11_OFF[r][j][2];
      dd[r] = LQ11_DIR[r][j];
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int x1 = tile % e1, q = tile / e1, x2 = q % e2, x3 = q / e2;
      const int tb = it & 1, k = it >> 1;
      if (lane == 0) {
        if (k >= 1) lq_mbar_wait_wd(&ownempty[tb], (unsigned)(k - 1) & 1u);
        lq_mbar_expect_tx(&ownfull[tb], 4 * 4608);
        lq_bulk_g2s(own + tb * 4 * LQ11_BLK, U + (lq_i64)tile * 36 * 32, 4 * 4608, &ownfull[tb]);
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        int y1 = x1 + o1[r], y2 = x2 + o2[r], y3 = x3 + o3[r];
        y1 = y1 < 0 ? y1 + e1 : (y1 >= e1 ? y1 - e1 : y1);
        y2 = y2 < 0 ? y2 + e2 : (y2 >= e2 ? y2 - e2 : y2);
        y3 = y3 < 0 ? y3 + e3 : (y3 >= e3 ? y3 - e3 : y3);
        const lq_i64 ch = y1 + (lq_i64)e1 * (y2 + (lq_i64)e2 * y3);
        if (lane == 0) {
          if (it >= 1) lq_mbar_wait_wd(&empty[r], (unsigned)(it - 1) & 1u);
          lq_mbar_expect_tx(&full[r], 11 * 4608);
        }
        __syncwarp();
        if (lane < 11) lq_bulk_g2s(stage + (r * 11 + lane) * LQ11_BLK, U + (ch * 36 + dd[r] * 9) * 32, 4608, &full[r]);
      }
    }
    return;
  }
  // ---------------- consumers: warp = mu + 4 * half (half 0: up staples + finish, half 1: down staples)
  const int mu = warp & 3, half = warp >> 2;
  const int x0 = lane < 16 ? 2 * lane : 2 * (lane - 16) + 1;
  const int x0p = (x0 + 1) & 31, x0m = (x0 + 31) & 31;
  const int lane_up = (x0p & 1) * 16 + (x0p >> 1), lane_dn = (x0m & 1) * 16 + (x0m >> 1);
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int tb = it & 1;
    cx ev[4];
    const int ee = (tile * 16 + mu * 4) * 32 + lane;
    if (half == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) ev[k] = __ldcs(E + ee + k * 32);
    }
    lq_mbar_wait_wd(&ownfull[tb], (unsigned)(it >> 1) & 1u);
    const cx* ownb = own + tb * 4 * LQ11_BLK;
    M3 acc = m3_zero();
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      lq_mbar_wait_wd(&full[r], (unsigned)it & 1u);
      const cx* sb = stage + r * 11 * LQ11_BLK;
      M3 op[3];
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        const int code = LQ11_OP[r][mu][half * 3 + o], sh = LQ11_SH[r][mu][half * 3 + o];
        const cx* blk = code < 16 ? sb + code * LQ11_BLK : ownb + (code - 16) * LQ11_BLK;
        op[o] = lq11_lds(blk, sh == 0 ? lane : (sh > 0 ? lane_up : lane_dn));
      }
      if (half == 0) {
        const M3 t = m3_mul_nd(op[0], op[1]);
        m3_fma_nd(acc, t, op[2]);
      } else {
        const M3 t = m3_mul_nn(op[0], op[1]);
        m3_fma_dn(acc, t, op[2]);
      }
      __syncwarp();
      if (lane == 0) lq_mbar_arrive(&empty[r]);
    }
    cx* pb = part + mu * LQ11_BLK;
    if (half == 1) {
      if (it >= 1) lq_mbar_wait_wd(&partempty[mu], (unsigned)(it - 1) & 1u);
#pragma unroll
      for (int k = 0; k < 9; ++k) pb[k * 32 + lane] = acc.e[k];
      __syncwarp();
      if (lane == 0) {
        lq_mbar_arrive(&partfull[mu]);
        lq_mbar_arrive(&ownempty[tb]);
      }
      continue;
    }
    const M3 u = lq11_lds(ownb + mu * LQ11_BLK, lane);
    __syncwarp();
    if (lane == 0) lq_mbar_arrive(&ownempty[tb]);
    lq_mbar_wait_wd(&partfull[mu], (unsigned)it & 1u);
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const cx d = pb[k * 32 + lane];
      acc.e[k] = cmk(acc.e[k].x + d.x, acc.e[k].y + d.y);
    }
    __syncwarp();
    if (lane == 0) lq_mbar_arrive(&partempty[mu]);
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
      cx* bo = Unew + (tile * 36 + mu * 9) * 32 + lane;
#pragma unroll
      for (int k = 0; k < 9; ++k) __stcs(bo + k * 32, un.e[k]);
    }
  }
}
