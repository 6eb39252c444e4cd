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
  const lq_i64 p = lq_phys(g, st.s);
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
    p = lq_phys(g, st.s);
    const Site<4> xpm = lq_up<4>(g, st, mu);
    {
      const Site<4> xpn = lq_up<4>(g, st, nu);
      M3 a = lq_load_link(U, g, nu, lq_phys(g, xpm.s));
      M3 b = lq_load_link(U, g, mu, lq_phys(g, xpn.s));
      M3 t = m3_mul_nd(a, b);
      M3 c = lq_load_link(U, g, nu, p);
      m3_fma_nd(acc, t, c);
    }
    {
      const Site<4> xmn = lq_dn<4>(g, st, nu);
      const Site<4> xpmmn = lq_dn<4>(g, xpm, nu);
      M3 a = lq_load_link(U, g, mu, lq_phys(g, xmn.s));
      M3 b = lq_load_link(U, g, nu, lq_phys(g, xpmmn.s));
      M3 t = m3_mul_nn(a, b);
      M3 c = lq_load_link(U, g, nu, lq_phys(g, xmn.s));
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
  const lq_i64 p = lq_phys(g, st.s);
  const Site<4> xpm = lq_up<4>(g, st, mu);
  const lq_i64 ppm = lq_phys(g, xpm.s);
  M3 acc = m3_zero();
#pragma unroll 1
  for (int j = 1; j < 4; ++j) {
    const int nu = (mu + j) & 3;
    const Site<4> xpn = lq_up<4>(g, st, nu);
    const Site<4> xmn = lq_dn<4>(g, st, nu);
    const Site<4> xpmmn = lq_dn<4>(g, xpm, nu);
    const lq_i64 pmn = lq_phys(g, xmn.s);
    {
      M3 a = lq_load_link(U, g, nu, ppm);
      M3 b = lq_load_link(U, g, mu, lq_phys(g, xpn.s));
      M3 t = m3_mul_nd(a, b);
      M3 c = lq_load_link(U, g, nu, p);
      m3_fma_nd(acc, t, c);
    }
    {
      M3 a = lq_load_link(U, g, mu, pmn);
      M3 b = lq_load_link(U, g, nu, lq_phys(g, xpmmn.s));
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

#endif  // !LQ_HOST_EMU
