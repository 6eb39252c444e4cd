// lq_tuned.cuh -- sm_100a-tuned D = 4 kernels of the molecular-dynamics hot loop (CUDA only).
//
// Same arithmetic as the generic functors KEfieldStep / KEfieldLinkStep of lq_kernels.cuh (the parity tests run
// both); what changes is the mapping of threads to links, the register budget and the memory-level parallelism.
// Variants that were measured and not adopted live in tools/lq_md_variants.cuh (kbench only).
#pragma once
#include <cstdlib>

#include "lq_kernels.cuh"

#ifndef LQ_HOST_EMU

// ------------------------------------------------------------------------------------------------------------
// Lean index arithmetic.  One thread per link, a warp = 32 consecutive site slots (one chunk of the chunked-SoA
// layout when ext0 is a multiple of 32) of one direction.  All neighbour slots are p + (sum of per-direction
// deltas): the eight deltas are computed once per thread, every matrix is one 32-bit element index -> IMAD.WIDE ->
// nine LDG.128 with immediate offsets.
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
// mbarrier / bulk-copy (TMA) primitives
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
// site decode of the per-link kernels (row walk, even x0 first) and the slot deltas of the eight neighbours
struct LqSite4 {
  int x0, x1, x2, x3;  // storage coordinates
  int p;               // slot
  int up[4], dn[4];    // slot deltas
};
// slot and neighbour deltas from the storage coordinates
__device__ __forceinline__ void lq_site4_fill(const LqGeom& g, LqSite4& s) {
  const int e0 = g.ext[0], ne0 = g.ne0;
  const int s1 = (int)g.sstride[1], s2 = (int)g.sstride[2], s3 = (int)g.sstride[3];
  const int sl0 = (s.x0 & 1) * ne0 + (s.x0 >> 1);
  s.p = s.x1 * s1 + s.x2 * s2 + s.x3 * s3 + sl0;
  const int x0p = s.x0 + 1 < e0 ? s.x0 + 1 : 0, x0m = s.x0 > 0 ? s.x0 - 1 : e0 - 1;
  s.up[0] = (x0p & 1) * ne0 + (x0p >> 1) - sl0;
  s.dn[0] = (x0m & 1) * ne0 + (x0m >> 1) - sl0;
  s.up[1] = s.x1 + 1 < g.sext[1] ? s1 : -s.x1 * s1;
  s.dn[1] = s.x1 > 0 ? -s1 : (g.sext[1] - 1) * s1;
  s.up[2] = s.x2 + 1 < g.sext[2] ? s2 : -s.x2 * s2;
  s.dn[2] = s.x2 > 0 ? -s2 : (g.sext[2] - 1) * s2;
  s.up[3] = s.x3 + 1 < g.sext[3] ? s3 : -s.x3 * s3;
  s.dn[3] = s.x3 > 0 ? -s3 : (g.sext[3] - 1) * s3;
}
__device__ __forceinline__ LqSite4 lq_site4(const LqGeom& g, int n) {
  LqSite4 s;
  const int e0 = g.ext[0], ne0 = g.ne0;
  int row = n / e0;
  const int lane = n - row * e0;
  s.x0 = lane < ne0 ? 2 * lane : 2 * (lane - ne0) + 1;
  int q = row / g.ext[1];
  s.x1 = row - q * g.ext[1] + g.ghost[1];
  row = q;
  q = row / g.ext[2];
  s.x2 = row - q * g.ext[2] + g.ghost[2];
  s.x3 = q + g.ghost[3];
  lq_site4_fill(g, s);
  return s;
}
// n-th site of colour `parity` (global coordinate sum & 1; even extents); gi = global reference-order SITE index
// (RNG stream ids are gi * 4 + mu: results do not depend on the decomposition)
__device__ __forceinline__ LqSite4 lq_site4_eo(const LqGeom& g, int n, int parity, lq_i64& gi) {
  LqSite4 s;
  const int h0 = g.ext[0] >> 1;
  int row = n / h0;
  const int k = n - row * h0;
  int q = row / g.ext[1];
  const int i1 = row - q * g.ext[1];
  row = q;
  q = row / g.ext[2];
  const int i2 = row - q * g.ext[2];
  const int i3 = q;
  s.x0 = 2 * k + ((parity + i1 + g.goff[1] + i2 + g.goff[2] + i3 + g.goff[3] + g.goff[0]) & 1);
  s.x1 = i1 + g.ghost[1];
  s.x2 = i2 + g.ghost[2];
  s.x3 = i3 + g.ghost[3];
  gi = (lq_i64)(s.x0 + g.goff[0]) * g.gstride[0] + (lq_i64)(i1 + g.goff[1]) * g.gstride[1] +
       (lq_i64)(i2 + g.goff[2]) * g.gstride[2] + (lq_i64)(i3 + g.goff[3]) * g.gstride[3];
  lq_site4_fill(g, s);
  return s;
}
// Staple sum around link (x, mu) in the accumulation order of lq_staple_sum (nu ascending, up then down: same bits as
// the generic functors), as the straight-line software pipeline of lq_md4_body: every operand is requested one
// product ahead of its first use.  `u` returns the link itself, read through the coherent path (the sweeps update it
// in place) while the last product runs.
__device__ __forceinline__ void lq_staples4(const cx* __restrict__ U, const cx* own, const LqSite4& s, int mu, M3& acc,
                                            M3& u) {
  const int p = s.p;
  const int pm = p + lq_sel4(mu, s.up[0], s.up[1], s.up[2], s.up[3]);
  const int n1 = mu == 0 ? 1 : 0, n2 = mu <= 1 ? 2 : 1, n3 = mu <= 2 ? 3 : 2;
  const int u1 = lq_sel4(n1, s.up[0], s.up[1], s.up[2], s.up[3]), d1 = lq_sel4(n1, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
  const int u2 = lq_sel4(n2, s.up[0], s.up[1], s.up[2], s.up[3]), d2 = lq_sel4(n2, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
  const int u3 = lq_sel4(n3, s.up[0], s.up[1], s.up[2], s.up[3]), d3 = lq_sel4(n3, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
  acc = m3_zero();
  M3 a = lq_ld36(U, pm, n1);
  M3 b = lq_ld36(U, p + u1, mu);
  M3 c, t;
#define LQ_STAGE_UP(NU, DN)      \
  c = lq_ld36(U, p, NU);         \
  t = m3_mul_nd(a, b);           \
  a = lq_ld36(U, p + DN, mu);    \
  b = lq_ld36(U, pm + DN, NU);   \
  m3_fma_nd(acc, t, c);
#define LQ_STAGE_DN(NU, DN, NEXTA, NEXTB) \
  c = lq_ld36(U, p + DN, NU);    \
  t = m3_mul_nn(a, b);           \
  a = NEXTA;                     \
  b = NEXTB;                     \
  m3_fma_dn(acc, t, c);
  LQ_STAGE_UP(n1, d1)
  LQ_STAGE_DN(n1, d1, lq_ld36(U, pm, n2), lq_ld36(U, p + u2, mu))
  LQ_STAGE_UP(n2, d2)
  LQ_STAGE_DN(n2, d2, lq_ld36(U, pm, n3), lq_ld36(U, p + u3, mu))
  LQ_STAGE_UP(n3, d3)
  c = lq_ld36(U, p + d3, n3);
  t = m3_mul_nn(a, b);
#pragma unroll
  for (int k = 0; k < 9; ++k) u.e[k] = own[k * 32];
  m3_fma_dn(acc, t, c);
#undef LQ_STAGE_UP
#undef LQ_STAGE_DN
}
// new boundary link / Gauss-field element -> the ghost layers of the neighbour ranks (peer memory over NVLink)
template <int NPL, int NV>
__device__ __forceinline__ void lq_push4(const LqGeom& g, const LqPush* __restrict__ ps, int x2, int x3, int p, int plane0,
                                         const cx* v) {
  const int o2 = g.ghost[2] ? (x2 == 1 ? 0 : (x2 == g.ext[2] ? 2 : 1)) : 1;
  const int o3 = g.ghost[3] ? (x3 == 1 ? 0 : (x3 == g.ext[3] ? 2 : 1)) : 1;
  if (o2 == 1 && o3 == 1) return;
#pragma unroll
  for (int c = 0; c < 3; ++c) {  // z-face, t-face, zt-corner neighbour
    const int a = c == 1 ? 1 : o2, bb = c == 0 ? 1 : o3;
    if ((a == 1 && bb == 1) || (c == 2 && (o2 == 1 || o3 == 1))) continue;
    const int k = ps->nbmap[a][bb];
    if (k < 0) continue;
    const int pd = p + ps->delta[k];
    cx* d = ps->peer[k] + ((pd >> 5) * NPL + plane0) * 32 + (pd & 31);
#pragma unroll
    for (int kk = 0; kk < NV; ++kk) d[kk * 32] = v[kk];
  }
}
// ------------------------------------------------------------------------------------------------------------
// Halo synchronisation folded into the compute kernels of a decomposed context (PUSH == 2).
// A ghost refresh used to be  kernel (compute + push) -> barrier kernel (release my epoch to the neighbours, acquire
// theirs) -> next kernel: 205 barrier launches per trajectory, each waiting for the slowest of the ranks.  Here the
// producing kernel releases the epoch itself -- the LAST of its boundary blocks to finish (a device counter), after a
// system-scope fence behind the pushes -- and the boundary blocks of the NEXT kernel acquire it before they touch a
// ghost layer.  Only blocks of boundary columns (x2 or x3 on a split face) read ghosts or push, and the grid visits them
// first, interleaved 1 : S with interior blocks: the epoch leaves after the first part of a kernel and is needed at the
// start of the neighbours' next one, so a rank may run ahead of its neighbours by most of a kernel before it waits.
// MEASURED ON 8 B200s (2 x 4 grid, 32^4 per GPU; profiles/r02r ... r02u, A/B on the same box) AND NOT THE DEFAULT: the
// barrier launches disappear (unaccounted time per trajectory 6.0 -> 1.5 ms) but the MD kernel takes 0.805 - 0.84 ms on
// every rank where the slowest rank's own compute is 0.745 ms (per-block system-scope fences behind NVLink stores and
// acquire loads in 3968 boundary blocks per launch), 127 - 130 ms per trajectory against 122 - 124 ms with barrier
// kernels; folding only the projection loop: 124.7 vs 122.2 ms.  Opt-in through LQ_FLAG_FOLD_HALO_SYNC.
// One acquire covers both hazards of the ping-pong buffers: the neighbour's pushes into my ghosts have landed (RAW),
// and its boundary blocks are done reading the ghosts this kernel's pushes overwrite (WAR).
struct LqFold {
  unsigned long long* remote[8];  // my slot in neighbour k's flag array (peer memory)
  unsigned long long* mine;       // my flag array: slot k = neighbour k's epoch
  unsigned int* counter;          // boundary blocks finished (zero between launches)
  unsigned long long wait_value;  // epoch the boundary blocks acquire first (0: none)
  unsigned long long signal_value;  // epoch released when the last boundary block is done (0: none)
  int n;    // neighbours
  int nB;   // boundary blocks of this launch (the last of them to finish releases the epoch)
  int nF;   // boundary blocks scheduled first, one every S grid positions
  int S;
  int bpc;  // blocks per (x2, x3) column
  int zfirst;  // 1: the z-face columns of the inner t-slices are scheduled first as well
};
// grid position -> natural block index (sites blk * BS ... in the reference order); isb: block of a boundary column.
// The first f.nF blocks of the schedule (one every S grid positions) are boundary blocks: all of them (zfirst), or those
// of the two boundary t-slices only -- the z faces then stay at their natural place in the sweep over x3, which keeps
// ONE stream of x3 planes in L2 (pulling the z faces of all 30 inner slices forward cost the MD kernel 8 % on 8 GPUs).
__device__ __forceinline__ int lq_fold_block(const LqGeom& g, const LqFold& f, int blk, bool& isb) {
  const int e2 = g.ext[2], e3 = g.ext[3];
  const int n3 = g.ghost[3] ? (e3 >= 2 ? 2 : 1) : 0;  // boundary values of x3, x2
  const int n2 = g.ghost[2] ? (e2 >= 2 ? 2 : 1) : 0;
  const int j = blk / f.S;
  int col, within;
  if (blk - j * f.S == 0 && j < f.nF) {
    isb = true;
    const int cb = j / f.bpc;
    within = j - cb * f.bpc;
    const int p1 = n3 * e2;  // columns of the boundary t-slices come first, then (zfirst) the z faces of the other slices
    int x2, x3;
    if (cb < p1) {
      const int q = cb / e2;
      x3 = q == 0 ? 0 : e3 - 1;
      x2 = cb - q * e2;
    } else {
      const int r = cb - p1, q = r / n2;
      x3 = (n3 ? 1 : 0) + q;
      x2 = r - q * n2 == 0 ? 0 : e2 - 1;
    }
    col = x2 + e2 * x3;
  } else {
    const int ji = blk - min((blk + f.S - 1) / f.S, f.nF);
    const int ci = ji / f.bpc;
    within = ji - ci * f.bpc;
    if (f.zfirst) {  // the rest: interior columns
      isb = false;
      const int i2 = e2 - n2, q = ci / i2;
      col = (n2 ? 1 : 0) + (ci - q * i2) + e2 * ((n3 ? 1 : 0) + q);
    } else {  // the rest: every column of the inner t-slices, z faces included
      const int q = ci / e2, x2 = ci - q * e2;
      isb = n2 && (x2 == 0 || x2 == e2 - 1);
      col = x2 + e2 * ((n3 ? 1 : 0) + q);
    }
  }
  return col * f.bpc + within;
}
__device__ __forceinline__ void lq_fold_wait(const LqFold& f) {  // all threads of a boundary block
  if (f.wait_value == 0) return;
  if ((int)threadIdx.x < f.n) {
    const long long t0 = clock64();
    unsigned long long v;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f.mine + threadIdx.x) : "memory");
      if (v >= f.wait_value) break;
      if (*(volatile unsigned long long*)(f.mine + 8) != 0) break;  // an earlier wait gave up: do not pile up timeouts
      if (clock64() - t0 > 20000000000ll) {  // ~10 s
        f.mine[8] = f.wait_value;  // error latch (LQ_P2P_MAXNB), reported by the next reduction / lq_sync
        break;
      }
      __nanosleep(100);
    }
  }
  __syncthreads();
}
__device__ __forceinline__ void lq_fold_signal(const LqFold& f) {  // all threads of a boundary block, after their pushes
  if (f.signal_value == 0) return;
  // one system-scope fence per block, behind the barrier that orders the other threads' pushes before it (a fence in
  // every thread made each boundary block wait for its own NVLink round trips: +0.09 ms per MD launch on 8 GPUs)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (atomicAdd(f.counter, 1u) + 1u == (unsigned)f.nB) {
      atomicExch(f.counter, 0u);
      __threadfence_system();
#pragma unroll
      for (int k = 0; k < 8; ++k)  // static indices: the parameter struct stays in the constant bank
        if (k < f.n) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f.remote[k]), "l"(f.signal_value) : "memory");
    }
  }
}
// Fused force + E kick (+ link step into Unew when FUSED; USE_EXP: U <- exp(i dt E) U instead of the Euler rule).
// The staple sum is a chain of twelve 3x3 products,  t = a b^+, acc += t c^+  (up)  and  t = a b, acc += t^+ c  (down)
// for nu ascending (the order of lq_staple_sum: same bits as the generic functor), written as STRAIGHT-LINE code with
// every operand requested one product ahead of its first use: ptxas then software-pipelines the whole chain (loads of
// product k+1 interleaved with the DFMAs of product k) inside the 168-register budget, where the rolled nu loop of
// round 1 exposed the full load latency at the top of each of its three iterations (0.654 -> 0.60 ms at 32^4;
// tools/kbench2.cu: v7; the variants that lost -- 128 / 96 register builds, two staples ahead, products fenced into
// basic blocks, three threads per link -- are in tools/lq_md_variants.cuh with their numbers in profiles/r02c).
// E and U' use streaming (evict-first) accesses.  PUSH: boundary links also go into the neighbour ranks' ghost layers.
template <int BLOCK, int FUSED, int USE_EXP, int PUSH>
__device__ __forceinline__ void lq_md4_body(const LqGeom& g, const cx* __restrict__ U, cx* __restrict__ Unew,
                                            cx* __restrict__ E, double coef, double dt_e, double dt_u, double c_u,
                                            int nkick, const LqPush* __restrict__ ps, int blk) {
  constexpr int SITES = BLOCK / 4;
  const int mu = threadIdx.x / SITES;
  const int n = blk * SITES + (threadIdx.x - mu * SITES);
  if (n >= (int)g.vol) return;
  const LqSite4 s = lq_site4(g, n);
  const int p = s.p;
  const int pm = p + lq_sel4(mu, s.up[0], s.up[1], s.up[2], s.up[3]);
  const int ee = ((p >> 5) * 16 + mu * 4) * 32 + (p & 31);
  // nu ascending with the own direction skipped
  const int n1 = mu == 0 ? 1 : 0, n2 = mu <= 1 ? 2 : 1, n3 = mu <= 2 ? 3 : 2;
  const int u1 = lq_sel4(n1, s.up[0], s.up[1], s.up[2], s.up[3]), d1 = lq_sel4(n1, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
  const int u2 = lq_sel4(n2, s.up[0], s.up[1], s.up[2], s.up[3]), d2 = lq_sel4(n2, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
  const int u3 = lq_sel4(n3, s.up[0], s.up[1], s.up[2], s.up[3]), d3 = lq_sel4(n3, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
  M3 acc = m3_zero();
  M3 a = lq_ld36(U, pm, n1);
  M3 b = lq_ld36(U, p + u1, mu);
  M3 c, t;
#define LQ_STAGE_UP(NU, DN)      /* up:  U_nu(x+mu) U_mu^+(x+nu) U_nu^+(x); then the (a, b) of the down staple */ \
  c = lq_ld36(U, p, NU);         \
  t = m3_mul_nd(a, b);           \
  a = lq_ld36(U, p + DN, mu);    \
  b = lq_ld36(U, pm + DN, NU);   \
  m3_fma_nd(acc, t, c);
#define LQ_STAGE_DN(NU, DN, NEXTA, NEXTB) /* down:  (U_mu(x-nu) U_nu(x+mu-nu))^+ U_nu(x-nu) */ \
  c = lq_ld36(U, p + DN, NU);    \
  t = m3_mul_nn(a, b);           \
  a = NEXTA;                     \
  b = NEXTB;                     \
  m3_fma_dn(acc, t, c);
  LQ_STAGE_UP(n1, d1)
  LQ_STAGE_DN(n1, d1, lq_ld36(U, pm, n2), lq_ld36(U, p + u2, mu))
  LQ_STAGE_UP(n2, d2)
  LQ_STAGE_DN(n2, d2, lq_ld36(U, pm, n3), lq_ld36(U, p + u3, mu))
  LQ_STAGE_UP(n3, d3)
  cx ev[4];
  c = lq_ld36(U, p + d3, n3);
  t = m3_mul_nn(a, b);
  a = lq_ld36(U, p, mu);  // the link itself, for U * A
#pragma unroll
  for (int k = 0; k < 4; ++k) ev[k] = __ldcs(E + ee + k * 32);
  m3_fma_dn(acc, t, c);
#undef LQ_STAGE_UP
#undef LQ_STAGE_DN
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
    M3 un = lq_link_update<4>(u, e, dt_u, c_u, USE_EXP);
    cx* bo = Unew + ((p >> 5) * 36 + mu * 9) * 32 + (p & 31);
#pragma unroll
    for (int k = 0; k < 9; ++k) __stcs(bo + k * 32, un.e[k]);
    if (PUSH) lq_push4<36, 9>(g, ps, s.x2, s.x3, p, mu * 9, un.e);
  }
}

template <int BLOCK, int MINB, int FUSED, int USE_EXP = 0, int PUSH = 0>
__global__ void __launch_bounds__(BLOCK, MINB)
    lq_md4_kernel(LqGeom g, const cx* __restrict__ U, cx* __restrict__ Unew, cx* __restrict__ E, double coef, double dt_e,
                  double dt_u, double c_u, int nkick, const LqPush* __restrict__ ps, int bps, const LqFold fold) {
  // ps: device-resident peer table, read by the threads of boundary slices only; bps: blocks per t-slice (0: keep
  // the natural block order); PUSH == 2: boundary columns first, halo synchronisation inside the kernel (LqFold)
  int blk = blockIdx.x;
  if (PUSH == 2) {
    bool isb;
    blk = lq_fold_block(g, fold, blk, isb);
    if (isb) lq_fold_wait(fold);
    lq_md4_body<BLOCK, FUSED, USE_EXP, 1>(g, U, Unew, E, coef, dt_e, dt_u, c_u, nkick, ps, blk);
    if (isb) lq_fold_signal(fold);
    return;
  }
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
  lq_md4_body<BLOCK, FUSED, USE_EXP, PUSH>(g, U, Unew, E, coef, dt_e, dt_u, c_u, nkick, ps, blk);
}
// ------------------------------------------------------------------------------------------------------------
// Checkerboard sweep sub-step (one direction mu, one colour), D = 4, with the lean addressing of V4: the staple sum
// of KHeatBath / KOverrelax (same accumulation order: nu ascending, up then down => the same bits as the generic
// functors) followed by the single-link rule.  KIND: 0 heat bath (heat_bath.rs:73-123), 1 over-relaxation
// (overrelaxation.rs:86-110, 158-184).  Links of the updated (mu, colour) set never enter each other's staples, so
// neighbours are read through the non-coherent path while the own link is read and written in place.
template <int BLOCK, int MINB, int KIND>
__global__ void __launch_bounds__(BLOCK, MINB)
    lq_sweep4_kernel(LqGeom g, cx* __restrict__ U, int mu, int parity, int flags, int or_kind, double coupling,
                     unsigned long long seed, unsigned long long counter) {
  const int n = blockIdx.x * BLOCK + threadIdx.x;
  if (n >= (int)(g.vol >> 1)) return;
  lq_i64 gi;
  const LqSite4 s = lq_site4_eo(g, n, parity, gi);
  cx* own = U + ((s.p >> 5) * 36 + mu * 9) * 32 + (s.p & 31);
  M3 acc, u;
  lq_staples4(U, own, s, mu, acc, u);
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
  if (PUSH) lq_push4<9, 9>(g, ps, x2, x3, p, 0, acc.e);
}

// ------------------------------------------------------------------------------------------------------------
// Gauss projection loop without the Gauss-field round trip, D = 4 (project_to_gauss, field.rs:1265-1337).
// The two-pass iteration (lq_gfield4_kernel + lq_gstep4_kernel) reads the links twice per iteration -- 2208 B/site --
// although they do not change during the ~100 iterations of a projection.  Here every link carries, next to E_i(x),
// its transported field  T_i(x) = U_i^+(x) E_i(x) U_i(x)  (Hermitian: stored as 3 real + 3 complex numbers in five
// 16-byte planes), so that the Gauss field is a sum of stored quantities,
//     G(x) = sum_j [ E_j(x) - T_j(x - j) ]                                               (field.rs:1174-1195)
// and ONE kernel per iteration (one thread per site) forms G(x) and the four G(x + i) on the fly, applies the
// projection step to the four links of the site (lq_gauss_project_link, field.rs:1301-1337) and writes E' and
// T' = U^+ E' U: links read once, no G array -- 1728 B/site.  Same arithmetic and summation order as lq_gauss_site /
// lq_gauss_project_link; the only difference from the two-pass path is that the anti-Hermitian rounding noise of
// U^+ E U (1e-17 relative) is not carried (the lower triangle is the conjugate of the upper one).
// RES: also reduce the residual sum_x |Tr((sum_a T_a) G(x))| (gauss_sum_div, field.rs:1199-1220) of the INPUT state
// into per-block partial sums.  PUSH: E' and T' of boundary sites also go into the neighbour ranks' ghost layers.
#define LQ_TPL 5 /* planes per link of the transported field: (h00, h11) (h22, 0) h01 h02 h12 */
__device__ __forceinline__ M3 lq_herm_unpack(const cx* __restrict__ b) {
  const cx q0 = __ldg(b), q1 = __ldg(b + 32), h01 = __ldg(b + 64), h02 = __ldg(b + 96), h12 = __ldg(b + 128);
  M3 m;
  m.e[0] = cmk(q0.x, 0.0);
  m.e[4] = cmk(q0.y, 0.0);
  m.e[8] = cmk(q1.x, 0.0);
  m.e[1] = h01;
  m.e[2] = h02;
  m.e[5] = h12;
  m.e[3] = cconj(h01);
  m.e[6] = cconj(h02);
  m.e[7] = cconj(h12);
  return m;
}
__device__ __forceinline__ void lq_herm_pack(const M3& m, cx v[LQ_TPL]) {
  v[0] = cmk(m.e[0].x, m.e[4].x);
  v[1] = cmk(m.e[8].x, 0.0);
  v[2] = m.e[1];
  v[3] = m.e[2];
  v[4] = m.e[5];
}
// G at the site of slot p from the stored fields; d[j] = slot delta to the site one step back in direction j
__device__ __forceinline__ M3 lq_gauss_from_et(const cx* __restrict__ E, const cx* __restrict__ T, int p, int d0, int d1,
                                               int d2, int d3) {
  M3 acc = m3_zero();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const cx* eb = E + ((p >> 5) * 16 + j * 4) * 32 + (p & 31);
    A8 e;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const cx v = __ldg(eb + k * 32);
      e.e[2 * k] = v.x;
      e.e[2 * k + 1] = v.y;
    }
    acc = m3_add(acc, lq_adjoint_to_matrix(e));
    const int pj = p + (j == 0 ? d0 : j == 1 ? d1 : j == 2 ? d2 : d3);
    acc = m3_sub(acc, lq_herm_unpack(T + ((pj >> 5) * (4 * LQ_TPL) + j * LQ_TPL) * 32 + (pj & 31)));
  }
  return acc;
}
// T = U^+ E U for every link (start of a projection loop)
template <int BLOCK, int PUSH>
__global__ void __launch_bounds__(BLOCK)
    lq_gtinit4_kernel(LqGeom g, const cx* __restrict__ U, const cx* __restrict__ E, cx* __restrict__ T,
                      const LqPush* __restrict__ psT) {
  constexpr int SITES = BLOCK / 4;
  const int mu = threadIdx.x / SITES;
  const int n = blockIdx.x * SITES + (threadIdx.x - mu * SITES);
  if (n >= (int)g.vol) return;
  const LqSite4 s = lq_site4(g, n);
  const int p = s.p;
  const cx* eb = E + ((p >> 5) * 16 + mu * 4) * 32 + (p & 31);
  A8 e;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const cx v = __ldg(eb + k * 32);
    e.e[2 * k] = v.x;
    e.e[2 * k + 1] = v.y;
  }
  const M3 u = lq_ld36(U, p, mu);
  const M3 t = m3_mul_dn(u, lq_adjoint_to_matrix(e));  // U^+ E
  M3 r = m3_zero();
  m3_fma_nn(r, t, u);
  cx tv[LQ_TPL];
  lq_herm_pack(r, tv);
  cx* tb = T + ((p >> 5) * (4 * LQ_TPL) + mu * LQ_TPL) * 32 + (p & 31);
#pragma unroll
  for (int k = 0; k < LQ_TPL; ++k) tb[k * 32] = tv[k];
  if (PUSH) lq_push4<4 * LQ_TPL, LQ_TPL>(g, psT, s.x2, s.x3, p, mu * LQ_TPL, tv);
}
// SH: the Gauss fields G(x) the threads of a block form for their own sites are passed through shared memory to the
// threads that need them as G(x + 0) (same x0 row: always inside the block) and G(x + 1) (the next row: three rows of
// four) -- the same bits, formed once instead of twice: 63 of the 232 128-bit loads of a thread and their additions go.
// Needs ext0 = 32 (a row = a warp), ext1 and the volume multiples of the block (the launcher checks).
template <int BLOCK, int MINB, int PUSH, int RES, int UNR = 0, int SH = 0>
__global__ void __launch_bounds__(BLOCK, MINB)
    lq_gausst4_kernel(LqGeom g, const cx* __restrict__ U, const cx* __restrict__ Ein, const cx* __restrict__ Tin,
                      cx* __restrict__ Eout, cx* __restrict__ Tout, double* __restrict__ partial,
                      const LqPush* __restrict__ psE, const LqPush* __restrict__ psT, const LqFold fold) {
  int blk = blockIdx.x;
  bool isb = false;
  if (PUSH == 2) {  // boundary columns first, halo synchronisation inside the kernel (see LqFold)
    blk = lq_fold_block(g, fold, blk, isb);
    if (isb) lq_fold_wait(fold);
  }
  const int n = blk * BLOCK + threadIdx.x;
  double res = 0.0;
  __shared__ cx gsm[SH ? 9 * BLOCK : 1];
  if (n < (int)g.vol) {
    const LqSite4 s = lq_site4(g, n);
    const int p = s.p;
    const M3 gx = lq_gauss_from_et(Ein, Tin, p, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
    if (SH) {  // (the volume is a multiple of the block: every thread gets here)
#pragma unroll
      for (int k = 0; k < 9; ++k) gsm[k * BLOCK + threadIdx.x] = gx.e[k];
      __syncthreads();
    }
    if (RES) {
      cx tr[8];
      lq_trace_gen(gx, tr);
      double re = 0.0, im = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        re += tr[k].x;
        im += tr[k].y;
      }
      res = sqrt(re * re + im * im);
    }
#pragma unroll UNR ? 4 : 1
    for (int i = 0; i < 4; ++i) {
      const int upi = lq_sel4(i, s.up[0], s.up[1], s.up[2], s.up[3]);
      const int pp = p + upi;
      M3 gp;
      // thread of the block that holds x + i: the x0 neighbour inside the warp's row, or the same lane one row up
      int src = -1;
      if (SH && i == 0) {
        const int x0p = s.x0 + 1 < 32 ? s.x0 + 1 : 0;
        src = (threadIdx.x & ~31) + (x0p & 1) * 16 + (x0p >> 1);
      } else if (SH && i == 1 && threadIdx.x + 32 < BLOCK) {
        src = threadIdx.x + 32;
      }
      if (src >= 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) gp.e[k] = gsm[k * BLOCK + src];
      } else {
        // the site x + i steps back to x in direction i and has the deltas of x in the other directions
        gp = lq_gauss_from_et(Ein, Tin, pp, i == 0 ? -upi : s.dn[0], i == 1 ? -upi : s.dn[1], i == 2 ? -upi : s.dn[2],
                              i == 3 ? -upi : s.dn[3]);
      }
      const int ee = ((p >> 5) * 16 + i * 4) * 32 + (p & 31);
      A8 e;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const cx v = __ldg(Ein + ee + k * 32);
        e.e[2 * k] = v.x;
        e.e[2 * k + 1] = v.y;
      }
      const M3 u = lq_ld36(U, p, i);
      e = lq_gauss_project_link(u, gx, gp, e);
      cx ev[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ev[k] = cmk(e.e[2 * k], e.e[2 * k + 1]);
        __stcs(Eout + ee + k * 32, ev[k]);
      }
      const M3 t = m3_mul_dn(u, lq_adjoint_to_matrix(e));  // U^+ E'
      M3 r = m3_zero();
      m3_fma_nn(r, t, u);
      cx tv[LQ_TPL];
      lq_herm_pack(r, tv);
      cx* tb = Tout + ((p >> 5) * (4 * LQ_TPL) + i * LQ_TPL) * 32 + (p & 31);
#pragma unroll
      for (int k = 0; k < LQ_TPL; ++k) tb[k * 32] = tv[k];
      if (PUSH) {
        lq_push4<16, 4>(g, psE, s.x2, s.x3, p, i * 4, ev);
        lq_push4<4 * LQ_TPL, LQ_TPL>(g, psT, s.x2, s.x3, p, i * LQ_TPL, tv);
      }
    }
  }
  if (RES) {
    __shared__ double sm[BLOCK / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) res += __shfl_down_sync(0xffffffffu, res, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = res;
    __syncthreads();
    if (threadIdx.x == 0) {
      double x = 0.0;
#pragma unroll
      for (int w = 0; w < BLOCK / 32; ++w) x += sm[w];
      partial[blk] = x;  // natural block order: the final sum does not depend on the visiting order
    }
  }
  if (PUSH == 2 && isb) lq_fold_signal(fold);
}

// the same iteration with one thread per LINK (four times the threads; G(x) is formed by each of the four threads of
// a site): A/B variant of lq_gausst4_kernel, residual taken by the mu = 0 threads
template <int BLOCK, int MINB, int PUSH, int RES>
__global__ void __launch_bounds__(BLOCK, MINB)
    lq_gausst4_link_kernel(LqGeom g, const cx* __restrict__ U, const cx* __restrict__ Ein, const cx* __restrict__ Tin,
                           cx* __restrict__ Eout, cx* __restrict__ Tout, double* __restrict__ partial,
                           const LqPush* __restrict__ psE, const LqPush* __restrict__ psT) {
  constexpr int SITES = BLOCK / 4;
  const int i = threadIdx.x / SITES;
  const int n = blockIdx.x * SITES + (threadIdx.x - i * SITES);
  double res = 0.0;
  if (n < (int)g.vol) {
    const LqSite4 s = lq_site4(g, n);
    const int p = s.p;
    const M3 gx = lq_gauss_from_et(Ein, Tin, p, s.dn[0], s.dn[1], s.dn[2], s.dn[3]);
    if (RES && i == 0) {
      cx tr[8];
      lq_trace_gen(gx, tr);
      double re = 0.0, im = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        re += tr[k].x;
        im += tr[k].y;
      }
      res = sqrt(re * re + im * im);
    }
    const int upi = lq_sel4(i, s.up[0], s.up[1], s.up[2], s.up[3]);
    const int pp = p + upi;
    const M3 gp = lq_gauss_from_et(Ein, Tin, pp, i == 0 ? -upi : s.dn[0], i == 1 ? -upi : s.dn[1], i == 2 ? -upi : s.dn[2],
                                   i == 3 ? -upi : s.dn[3]);
    const int ee = ((p >> 5) * 16 + i * 4) * 32 + (p & 31);
    A8 e;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const cx v = __ldg(Ein + ee + k * 32);
      e.e[2 * k] = v.x;
      e.e[2 * k + 1] = v.y;
    }
    const M3 u = lq_ld36(U, p, i);
    e = lq_gauss_project_link(u, gx, gp, e);
    cx ev[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ev[k] = cmk(e.e[2 * k], e.e[2 * k + 1]);
      __stcs(Eout + ee + k * 32, ev[k]);
    }
    const M3 t = m3_mul_dn(u, lq_adjoint_to_matrix(e));
    M3 r = m3_zero();
    m3_fma_nn(r, t, u);
    cx tv[LQ_TPL];
    lq_herm_pack(r, tv);
    cx* tb = Tout + ((p >> 5) * (4 * LQ_TPL) + i * LQ_TPL) * 32 + (p & 31);
#pragma unroll
    for (int k = 0; k < LQ_TPL; ++k) tb[k * 32] = tv[k];
    if (PUSH) {
      lq_push4<16, 4>(g, psE, s.x2, s.x3, p, i * 4, ev);
      lq_push4<4 * LQ_TPL, LQ_TPL>(g, psT, s.x2, s.x3, p, i * LQ_TPL, tv);
    }
  }
  if (RES) {
    __shared__ double sm[BLOCK / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) res += __shfl_down_sync(0xffffffffu, res, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = res;
    __syncthreads();
    if (threadIdx.x == 0) {
      double x = 0.0;
#pragma unroll
      for (int w = 0; w < BLOCK / 32; ++w) x += sm[w];
      partial[blockIdx.x] = x;
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
    lq_i64 gi;
    const LqSite4 s = lq_site4_eo(g, n, parity, gi);
    cx* own = U + ((s.p >> 5) * 36 + mu * 9) * 32 + (s.p & 31);
    M3 acc, old;
    lq_staples4(U, own, s, mu, acc, old);
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
#ifndef LQ_TUNED_NO_LAUNCHERS  /* tools/ experiments include the kernels without instantiating the product set */
// launchers used by lq_capi.cu (D = 4; 32-bit element indices: fields below 2^31 elements, else the generic path)
static inline bool lq_tuned_ok(const LqGeom& g) {
  return g.D == 4 && g.nchunk * 32 * 36 < ((lq_i64)1 << 31) && g.vol < ((lq_i64)1 << 31);
}
static inline cudaError_t lq_tuned_efield_step(cudaStream_t st, const LqGeom& g, const cx* U, cx* E, double coef, double dt,
                                               int nkick) {
  constexpr int BLOCK = 128;
  lq_md4_kernel<BLOCK, 3, 0><<<(unsigned)((g.vol + BLOCK / 4 - 1) / (BLOCK / 4)), BLOCK, 0, st>>>(g, U, nullptr, E, coef, dt,
                                                                                                0.0, 0.0, nkick, nullptr, 0, LqFold{});
  return cudaGetLastError();
}
static inline cudaError_t lq_tuned_efield_link_step(cudaStream_t st, const LqGeom& g, const cx* U, cx* Unew, cx* E,
                                                    double coef, double dt_e, double dt_u, double c_u, int nkick,
                                                    int use_exp = 0) {
  constexpr int BLOCK = 128;
  const unsigned nb = (unsigned)((g.vol + BLOCK / 4 - 1) / (BLOCK / 4));
  if (use_exp)
    lq_md4_kernel<BLOCK, 3, 1, 1><<<nb, BLOCK, 0, st>>>(g, U, Unew, E, coef, dt_e, dt_u, c_u, nkick, nullptr, 0, LqFold{});
  else
    lq_md4_kernel<BLOCK, 3, 1, 0><<<nb, BLOCK, 0, st>>>(g, U, Unew, E, coef, dt_e, dt_u, c_u, nkick, nullptr, 0, LqFold{});
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
// projection loop on the transported field (lq_gausst4_kernel): T = U^+ E U, then one kernel per iteration
static inline size_t lq_tuned_t_bytes(const LqGeom& g) { return (size_t)g.nchunk * 32 * 4 * LQ_TPL * sizeof(cx); }
static inline cudaError_t lq_tuned_gauss_tinit(cudaStream_t st, const LqGeom& g, const cx* U, const cx* E, cx* T,
                                               const LqPush* psT) {
  constexpr int BLOCK = 128;
  const unsigned nb = (unsigned)((g.vol + BLOCK / 4 - 1) / (BLOCK / 4));
  if (psT) lq_gtinit4_kernel<BLOCK, 1><<<nb, BLOCK, 0, st>>>(g, U, E, T, psT);
  else lq_gtinit4_kernel<BLOCK, 0><<<nb, BLOCK, 0, st>>>(g, U, E, T, nullptr);
  return cudaGetLastError();
}
// variant: 0 = one thread per site, rolled loop over the four links; 1 = the same, unrolled; 2 = one thread per link
static inline lq_i64 lq_tuned_gausst_blocks(const LqGeom& g, int variant) {
  (void)variant;
  return (g.vol + 127) / 128;
}
// Gauss fields shared inside a block (template flag SH of lq_gausst4_kernel): a row must be a warp and a block four
// whole rows of one (x2, x3) column; LQ_GAUSST_NO_SHARE=1 in the environment keeps the unshared kernel (A/B)
static inline bool lq_tuned_gausst_share_ok(const LqGeom& g) {
  static const bool off = getenv("LQ_GAUSST_NO_SHARE") != nullptr;
  return !off && g.ext[0] == 32 && g.ext[1] % 4 == 0 && g.vol % 128 == 0 && !g.ghost[0] && !g.ghost[1];
}
template <int PUSH, int RES>
static inline void lq_tuned_gauss_titer_launch(cudaStream_t st, const LqGeom& g, const cx* U, const cx* Ein, const cx* Tin,
                                               cx* Eout, cx* Tout, double* partial, const LqPush* psE, const LqPush* psT,
                                               int variant) {
  constexpr int BLOCK = 128;
  const unsigned nb = (unsigned)lq_tuned_gausst_blocks(g, variant);
  if (variant == 2) lq_gausst4_kernel<BLOCK, 5, PUSH, RES, 0><<<nb, BLOCK, 0, st>>>(g, U, Ein, Tin, Eout, Tout, partial, psE, psT, LqFold{});
  else if (variant == 1) lq_gausst4_kernel<BLOCK, 4, PUSH, RES, 0><<<nb, BLOCK, 0, st>>>(g, U, Ein, Tin, Eout, Tout, partial, psE, psT, LqFold{});
  else if (variant == 3) lq_gausst4_kernel<BLOCK, 6, PUSH, RES, 0><<<nb, BLOCK, 0, st>>>(g, U, Ein, Tin, Eout, Tout, partial, psE, psT, LqFold{});
  else if (variant == 0 && lq_tuned_gausst_share_ok(g))
    lq_gausst4_kernel<BLOCK, 3, PUSH, RES, 0, 1><<<nb, BLOCK, 0, st>>>(g, U, Ein, Tin, Eout, Tout, partial, psE, psT, LqFold{});
  else lq_gausst4_kernel<BLOCK, 3, PUSH, RES, 0><<<nb, BLOCK, 0, st>>>(g, U, Ein, Tin, Eout, Tout, partial, psE, psT, LqFold{});
}
static inline cudaError_t lq_tuned_gauss_titer(cudaStream_t st, const LqGeom& g, const cx* U, const cx* Ein, const cx* Tin,
                                               cx* Eout, cx* Tout, double* partial, bool want_res, const LqPush* psE,
                                               const LqPush* psT, int variant) {
  if (psE) {
    if (want_res) lq_tuned_gauss_titer_launch<1, 1>(st, g, U, Ein, Tin, Eout, Tout, partial, psE, psT, variant);
    else lq_tuned_gauss_titer_launch<1, 0>(st, g, U, Ein, Tin, Eout, Tout, partial, psE, psT, variant);
  } else {
    if (want_res) lq_tuned_gauss_titer_launch<0, 1>(st, g, U, Ein, Tin, Eout, Tout, partial, nullptr, nullptr, variant);
    else lq_tuned_gauss_titer_launch<0, 0>(st, g, U, Ein, Tin, Eout, Tout, partial, nullptr, nullptr, variant);
  }
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
    lq_md4_kernel<BLOCK, 3, 1, 1, 1><<<nb, BLOCK, 0, st>>>(g, U, Unew, E, coef, dt_e, dt_u, c_u, nkick, d_ps, bps, LqFold{});
  else
    lq_md4_kernel<BLOCK, 3, 1, 0, 1><<<nb, BLOCK, 0, st>>>(g, U, Unew, E, coef, dt_e, dt_u, c_u, nkick, d_ps, bps, LqFold{});
  return cudaGetLastError();
}

// ---- launches with the halo synchronisation folded in (PUSH == 2).  lq_fold_geom fills the block schedule of a kernel
// whose blocks hold `bs` consecutive sites; false: the geometry is not covered (blocks would straddle (x2, x3) columns)
static inline bool lq_fold_geom(const LqGeom& g, int bs, int zfirst, LqFold& f) {
  if (g.D != 4 || (!g.ghost[2] && !g.ghost[3]) || g.ghost[0] || g.ghost[1]) return false;
  const lq_i64 colsites = (lq_i64)g.ext[0] * g.ext[1];
  if (colsites % bs != 0 || g.vol % bs != 0) return false;
  const int e2 = g.ext[2], e3 = g.ext[3];
  const int n3 = g.ghost[3] ? (e3 >= 2 ? 2 : 1) : 0, n2 = g.ghost[2] ? (e2 >= 2 ? 2 : 1) : 0;
  const lq_i64 colB = (lq_i64)n3 * e2 + (lq_i64)(e3 - n3) * n2;
  const lq_i64 bpc = colsites / bs, total = (lq_i64)e2 * e3 * bpc, nB = colB * bpc;
  if (nB <= 0 || nB > total || total > 0x7fffffff) return false;
  f.bpc = (int)bpc;
  f.nB = (int)nB;
  f.zfirst = (zfirst || n3 == 0) ? 1 : 0;  // no t split: the z faces are all there is to schedule first
  const lq_i64 nF = f.zfirst ? nB : (lq_i64)n3 * e2 * bpc;
  f.nF = (int)nF;
  const lq_i64 sp = total / nF;
  f.S = (int)(sp < 1 ? 1 : (sp > 4 ? 4 : sp));
  return true;
}
static inline cudaError_t lq_tuned_efield_link_step_fold(cudaStream_t st, const LqGeom& g, const cx* U, cx* Unew, cx* E,
                                                         double coef, double dt_e, double dt_u, double c_u, int nkick,
                                                         const LqPush* d_ps, int use_exp, const LqFold& fold) {
  constexpr int BLOCK = 128;
  const unsigned nb = (unsigned)(g.vol / (BLOCK / 4));
  if (use_exp)
    lq_md4_kernel<BLOCK, 3, 1, 1, 2><<<nb, BLOCK, 0, st>>>(g, U, Unew, E, coef, dt_e, dt_u, c_u, nkick, d_ps, 0, fold);
  else
    lq_md4_kernel<BLOCK, 3, 1, 0, 2><<<nb, BLOCK, 0, st>>>(g, U, Unew, E, coef, dt_e, dt_u, c_u, nkick, d_ps, 0, fold);
  return cudaGetLastError();
}
static inline cudaError_t lq_tuned_gauss_titer_fold(cudaStream_t st, const LqGeom& g, const cx* U, const cx* Ein,
                                                    const cx* Tin, cx* Eout, cx* Tout, double* partial, bool want_res,
                                                    const LqPush* psE, const LqPush* psT, const LqFold& fold) {
  constexpr int BLOCK = 128;
  const unsigned nb = (unsigned)(g.vol / BLOCK);
  const bool sh = lq_tuned_gausst_share_ok(g);
  if (want_res && sh)
    lq_gausst4_kernel<BLOCK, 3, 2, 1, 0, 1><<<nb, BLOCK, 0, st>>>(g, U, Ein, Tin, Eout, Tout, partial, psE, psT, fold);
  else if (want_res)
    lq_gausst4_kernel<BLOCK, 3, 2, 1, 0><<<nb, BLOCK, 0, st>>>(g, U, Ein, Tin, Eout, Tout, partial, psE, psT, fold);
  else if (sh)
    lq_gausst4_kernel<BLOCK, 3, 2, 0, 0, 1><<<nb, BLOCK, 0, st>>>(g, U, Ein, Tin, Eout, Tout, partial, psE, psT, fold);
  else
    lq_gausst4_kernel<BLOCK, 3, 2, 0, 0><<<nb, BLOCK, 0, st>>>(g, U, Ein, Tin, Eout, Tout, partial, psE, psT, fold);
  return cudaGetLastError();
}

#endif  // !LQ_TUNED_NO_LAUNCHERS

#endif  // !LQ_HOST_EMU
