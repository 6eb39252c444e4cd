// lq_local.cuh -- counter-based RNG and the single-link update rules of the local sweeps
// (heat bath, over-relaxation, Metropolis).  All host/device inline; one thread updates one link.
#pragma once
#include "lq_common.cuh"

// ---------------------------------------------------------------------------------------------- Philox4x32-10
// Salmon et al. SC'11.  One stream per (seed, call counter, global link index); 53-bit uniforms, two per block.
// (The reference draws from rand 0.8 StdRng, an un-vendored dependency whose stream no reference test pins;
//  the library and the oracle share this specified generator instead.)
struct LqU4 {
  uint32_t x, y, z, w;
};
// One Philox4x32-10 block.  Deliberately NOT inlined: the samplers draw at a dozen call sites per kernel and the ten
// rounds would otherwise be replicated at each of them (instruction-cache misses were 15 % of the stall samples of
// the heat-bath kernel).  Arguments and result travel in registers.
LQ_NOINLINE LqU4 lq_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  LqU4 r;
  r.x = c0;
  r.y = c1;
  r.z = c2;
  r.w = c3;
  return r;
}
struct LqStream {
  uint32_t key[2], ctr[4], buf[4];
  int have;
  LQ_HD LqStream(uint64_t seed, uint64_t counter, uint64_t idx) {
    key[0] = (uint32_t)seed;
    key[1] = (uint32_t)(seed >> 32);
    ctr[0] = 0;
    ctr[1] = (uint32_t)idx;
    ctr[2] = (uint32_t)counter;
    ctr[3] = ((uint32_t)(counter >> 32) & 0x00FFFFFFu) | (((uint32_t)(idx >> 32) & 0xFFu) << 24);
    have = 0;
    buf[0] = buf[1] = buf[2] = buf[3] = 0;
  }
  LQ_HD void block() {
    LqU4 r = lq_philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    buf[0] = r.x;
    buf[1] = r.y;
    buf[2] = r.z;
    buf[3] = r.w;
  }
  LQ_HD void words(uint32_t& hi, uint32_t& lo) {
    if (have == 0) {
      block();
      ctr[0] += 1;
      have = 2;
    }
    int o = (2 - have) * 2;
    --have;
    hi = o == 0 ? buf[0] : buf[2];
    lo = o == 0 ? buf[1] : buf[3];
  }
  // 53-bit uniform  ((hi >> 5) * 2^26 + (lo >> 6)) * 2^-53  from two exact 32-bit conversions and one FMA (27 + 26 bits:
  // every step is exact, so this IS the 64-bit integer conversion, without its multi-instruction device sequence)
  LQ_HD double uniform01() {  // [0,1)
    uint32_t hi, lo;
    words(hi, lo);
    return fma((double)(hi >> 5), 0x1.0p-27, (double)(lo >> 6) * 0x1.0p-53);
  }
  LQ_HD double open_closed01() { return uniform01() + 0x1.0p-53; }  // (0,1]: (bits + 1) * 2^-53, exactly
  LQ_HD double uniform_pm1() { return 2.0 * uniform01() - 1.0; }              // [-1,1)
  LQ_HD bool bernoulli(double p) { return uniform01() < p; }
  LQ_HD void normal_pair(double& z0, double& z1) {  // Box-Muller from one Philox block
    double u1 = open_closed01();
    double u2 = uniform01();
    double r = sqrt(-2.0 * log(u1));
    double th = 2.0 * LQ_PI * u2;
    z0 = r * cos(th);
    z1 = r * sin(th);
  }
};

// cos(2 pi x): on the device cospi has an exact argument reduction (no slow path); the host build keeps libm's cos
#if defined(__CUDA_ARCH__)
#define LQ_COS_2PI(x) cospi(2.0 * (x))
#else
#define LQ_COS_2PI(x) cos(2.0 * LQ_PI * (x))
#endif
// 1/sqrt(x): rsqrt is a CUDA function (also callable from host code under nvcc); the g++ host build has none
#ifdef LQ_HOST_EMU
#define LQ_RSQRT(x) (1.0 / sqrt(x))
#else
#define LQ_RSQRT(x) rsqrt(x)
#endif
#define LQ_KP_MAX_ITER 10000 /* the reference loops forever on NaN parameters; cap and return x0 = 1 */

// ---------------------------------------------------------------------------------------------- 2x2
struct M2 {
  cx a, b, c, d;  // [[a, b], [c, d]]
};
LQ_HD M2 m2_mul(const M2& x, const M2& y) {
  M2 r;
  r.a = cadd(cmul(x.a, y.a), cmul(x.b, y.c));
  r.b = cadd(cmul(x.a, y.b), cmul(x.b, y.d));
  r.c = cadd(cmul(x.c, y.a), cmul(x.d, y.c));
  r.d = cadd(cmul(x.c, y.b), cmul(x.d, y.d));
  return r;
}
LQ_HD M2 m2_adj(const M2& x) {
  M2 r;
  r.a = cconj(x.a);
  r.b = cconj(x.c);
  r.c = cconj(x.b);
  r.d = cconj(x.d);
  return r;
}
LQ_HD cx m2_det(const M2& x) { return csub(cmul(x.a, x.d), cmul(x.c, x.b)); }
LQ_HD bool lq_is_normal(double v) { return isfinite(v) && fabs(v) >= DBL_MIN; }

// complex_matrix_from_vec, su2.rs:134-140 -- PAULI_3 as coded is diag(1,1) (su2.rs:39-45) unless the
// context carries LQ_FLAG_PAULI3_FIXED (bit 0).
LQ_HD M2 lq_matrix_from_vec(double x0, const double x[3], int flags) {
  M2 r;
  r.a = cmk(x0, x[2]);
  r.b = cmk(x[1], x[0]);
  r.c = cmk(-x[1], x[0]);
  r.d = (flags & 1) ? cmk(x0, -x[2]) : cmk(x0, x[2]);
  return r;
}
// random_su2_close_to_unity, su2.rs:80-100
LQ_HD M2 lq_random_su2_close_to_unity(double spread, LqStream& rng, int flags) {
  double r[3];
  r[0] = rng.uniform_pm1();
  r[1] = rng.uniform_pm1();
  r[2] = rng.uniform_pm1();
  // try_normalize(eps) of su2.rs:88 with one reciprocal square root for the three components (<= 2 ulp from dividing)
  const double n2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  const double sc = (n2 > LQ_EPS * LQ_EPS ? LQ_RSQRT(n2) : 1.0) * spread;
  double x[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) x[k] = r[k] * sc;
  double x0u = sqrt(1.0 - (x[0] * x[0] + x[1] * x[1] + x[2] * x[2]));
  double x0 = rng.bernoulli(0.5) ? x0u : -x0u;
  return lq_matrix_from_vec(x0, x, flags);
}
// get_r / get_s / get_t (su3.rs:428-530): embed a 2x2 block into rows/cols (0,1), (0,2), (1,2)
LQ_HD M3 lq_embed(const M2& m, int which) {
  const int ia = which == 2 ? 1 : 0;
  const int ib = which == 0 ? 1 : 2;
  M3 r = m3_ident();
  r.e[3 * ia + ia] = m.a;
  r.e[3 * ia + ib] = m.b;
  r.e[3 * ib + ia] = m.c;
  r.e[3 * ib + ib] = m.d;
  return r;
}
// get_sub_block_{r,s,t}, su3.rs:559-620
LQ_HD M2 lq_sub_block(const M3& m, int which) {
  const int ia = which == 2 ? 1 : 0;
  const int ib = which == 0 ? 1 : 2;
  M2 r;
  r.a = m.e[3 * ia + ia];
  r.b = m.e[3 * ia + ib];
  r.c = m.e[3 * ib + ia];
  r.d = m.e[3 * ib + ib];
  return r;
}
// project_to_su2_unorm, su2.rs:155-157:  m - m^dagger + 1 * conj(tr m)
LQ_HD M2 lq_project_to_su2_unorm(const M2& m) {
  M2 ad = m2_adj(m), r;
  cx t = cconj(cadd(m.a, m.d));
  r.a = cadd(csub(m.a, ad.a), t);
  r.b = csub(m.b, ad.b);
  r.c = csub(m.c, ad.c);
  r.d = cadd(csub(m.d, ad.d), t);
  return r;
}
// random_su3_close_to_unity, su3.rs:385-398
LQ_HD M3 lq_random_su3_close_to_unity(double spread, LqStream& rng, int flags) {
  M3 r = lq_embed(lq_random_su2_close_to_unity(spread, rng, flags), 0);
  M3 s = lq_embed(lq_random_su2_close_to_unity(spread, rng, flags), 1);
  M3 t = lq_embed(lq_random_su2_close_to_unity(spread, rng, flags), 2);
  M3 x = m3_mul_nn(m3_mul_nn(r, s), t);
  if (rng.bernoulli(0.5)) x = m3_adj(x);
  return x;
}
// random_su3, su3.rs:322-355 (Gram-Schmidt of two Uniform(-1,1)^6 vectors)
LQ_HD M3 lq_random_su3(LqStream& rng) {
  cx v1[3], v2[3];
  int guard = 0;
  do {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double re = rng.uniform_pm1();
      double im = rng.uniform_pm1();
      v1[k] = cmk(re, im);
    }
  } while (sqrt(cnorm2(v1[0]) + cnorm2(v1[1]) + cnorm2(v1[2])) <= LQ_EPS && ++guard < LQ_KP_MAX_ITER);
  guard = 0;
  cx dot;
  do {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double re = rng.uniform_pm1();
      double im = rng.uniform_pm1();
      v2[k] = cmk(re, im);
    }
    dot = cadd(cadd(cmul(v1[0], v2[0]), cmul(v1[1], v2[1])), cmul(v1[2], v2[2]));  // non-conjugating, su3.rs:343
  } while (sqrt(cnorm2(dot)) <= LQ_EPS && ++guard < LQ_KP_MAX_ITER);
  M3 m = m3_zero();
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    m.e[3 * k + 0] = v1[k];
    m.e[3 * k + 1] = v2[k];
  }
  return lq_orthonormalize(m);
}
// random_su2, su2.rs:200-216
LQ_HD M2 lq_random_su2(LqStream& rng) {
  cx v0, v1;
  double n;
  int guard = 0;
  do {
    double re = rng.uniform_pm1();
    double im = rng.uniform_pm1();
    v0 = cmk(re, im);
    re = rng.uniform_pm1();
    im = rng.uniform_pm1();
    v1 = cmk(re, im);
    n = sqrt(cnorm2(v0) + cnorm2(v1));
  } while (!lq_is_normal(n) && ++guard < LQ_KP_MAX_ITER);
  cx a = cmk(v0.x / n, v0.y / n), b = cmk(v1.x / n, v1.y / n);
  M2 r;
  r.a = a;
  r.b = b;
  r.c = cneg(cconj(b));
  r.d = cconj(a);
  return r;
}
// ModifiedNormal + HeatBathDistributionNorm (Kennedy-Pendleton), distribution.rs:89-100, 336-350
// Only lambda^2 = -(ln r0 + cos^2(2 pi r1) ln r2) / (2 alpha) enters the accept test and the result, so the square root
// of distribution.rs:96 (and the squaring that follows it) is not taken, and the division is one reciprocal per call:
// an ulp or two away from the literal form, far inside the 1e-9 tolerance of the stochastic paths.
LQ_HD double lq_heat_bath_norm(double param_exp, LqStream& rng) {
  const double inv2a = 0.5 / param_exp;
  for (int it = 0; it < LQ_KP_MAX_ITER; ++it) {
    double r = rng.uniform01();
    double r0 = rng.open_closed01();
    double r1 = rng.open_closed01();
    double r2 = rng.open_closed01();
    double c = LQ_COS_2PI(r1);
    double l2 = -(log(r0) + c * c * log(r2)) * inv2a;
    if (r * r <= 1.0 - l2) return 1.0 - 2.0 * l2;
  }
  return 1.0;
}
// HeatBathDistribution -> 2x2 matrix, distribution.rs:199-219 (direction = normalised cube sample, as coded; with
// LQ_FLAG_UNIFORM_DIRECTION cube samples outside the unit ball are rejected => uniform on the sphere)
LQ_HD M2 lq_heat_bath_matrix(double param_exp, LqStream& rng, int flags) {
  double x0 = lq_heat_bath_norm(param_exp, rng);
  double xu[3], n2;
  int guard = 0;
  do {
    xu[0] = rng.uniform_pm1();
    xu[1] = rng.uniform_pm1();
    xu[2] = rng.uniform_pm1();
    n2 = xu[0] * xu[0] + xu[1] * xu[1] + xu[2] * xu[2];  // the tests of distribution.rs:209-213 on the squared length
  } while ((n2 <= LQ_EPS * LQ_EPS || ((flags & 16) && n2 > 1.0)) && ++guard < LQ_KP_MAX_ITER);  // 16: LQ_FLAG_UNIFORM_DIRECTION
  // x = xu / |xu| * sqrt(1 - x0^2) with one reciprocal square root instead of a square root and three divisions (the
  // reference divides each component, distribution.rs:214: the results differ by an ulp or two, far inside the 1e-9
  // tolerance of the stochastic paths)
  const double sc = sqrt(1.0 - x0 * x0) * LQ_RSQRT(n2);
  double x[3] = {xu[0] * sc, xu[1] * sc, xu[2] * sc};
  return lq_matrix_from_vec(x0, x, flags);
}
// heat_bath_su2, heat_bath.rs:73-86.  coupling = beta * coupling_scale (reference: scale 1, i.e. beta*k).
LQ_HD M2 lq_heat_bath_su2(const M2& stap, double coupling, LqStream& rng, int flags) {
  double k = sqrt(m2_det(stap).x);
  if (lq_is_normal(k)) {
    M2 v = m2_adj(stap);
    const double rk = 1.0 / k;  // one reciprocal instead of eight divisions (heat_bath.rs:80 divides; <= 1 ulp apart)
    v.a = cmk(v.a.x * rk, v.a.y * rk);
    v.b = cmk(v.b.x * rk, v.b.y * rk);
    v.c = cmk(v.c.x * rk, v.c.y * rk);
    v.d = cmk(v.d.x * rk, v.d.y * rk);
    M2 x = lq_heat_bath_matrix(coupling * k, rng, flags);
    return m2_mul(x, v);
  }
  return lq_random_su2(rng);
}
// HeatBathSweep::get_modif, heat_bath.rs:90-109: Cabibbo-Marinari over the r, s, t SU(2) blocks:
//   r = R(U A), s = S(r U A), t = T(s r U A), U' = t s r U -- one loop body executed three times (a third of the
//   code of the unrolled form: the sampler with its Philox rounds, log and cos is instantiated once).
// Only what the rule consumes is computed: the 2x2 block (ia, ib) of W = cur * A (4 of its 9 entries), and the two
// rows (ia, ib) of x * cur that the embedded 2x2 matrix x changes.  Every entry is the same k-ascending FMA chain as in
// m3_mul_nn -- the skipped terms are products with the exact 0 / 1 entries of the embedding -- so the bits are those of
// the two full 3x3 products (44 % of their FMAs).
// The staple sum `a` enters through an accessor: registers (LqStapleRegs) or, for the warp-specialised sweep kernel whose
// rule warps run on a small register budget, the shared-memory slot the staple warps filled (LqStapleShared).
struct LqStapleRegs {
  const M3& a;
  LQ_HD void cols(int which, cx ca[3], cx cb[3]) const {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      ca[k] = which == 2 ? a.e[3 * k + 1] : a.e[3 * k];
      cb[k] = which == 0 ? a.e[3 * k + 1] : a.e[3 * k + 2];
    }
  }
};
template <int STRIDE>
struct LqStapleShared {
  const cx* a;  // entry k of the 3x3 staple sum at a[k * STRIDE]
  LQ_HD void cols(int which, cx ca[3], cx cb[3]) const {
    const int c0 = which == 2 ? 1 : 0, c1 = which == 0 ? 1 : 2;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      ca[k] = a[(3 * k + c0) * STRIDE];
      cb[k] = a[(3 * k + c1) * STRIDE];
    }
  }
};
template <class Staple, class Rule>
LQ_HD M3 lq_subgroup_update_acc(const M3& u, const Staple& a, Rule rule) {
  M3 cur = u;
#pragma unroll 1
  for (int which = 0; which < 3; ++which) {
    // blocks (0,1), (0,2), (1,2): ia = which == 2, ib = 1 + (which != 0)
    cx ra[3], rb[3], ca[3], cb[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      ra[k] = which == 2 ? cur.e[3 + k] : cur.e[k];
      rb[k] = which == 0 ? cur.e[3 + k] : cur.e[6 + k];
    }
    a.cols(which, ca, cb);
    M2 w;
    w.a = w.b = w.c = w.d = cmk(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      cfma(w.a, ra[k], ca[k]);
      cfma(w.b, ra[k], cb[k]);
      cfma(w.c, rb[k], ca[k]);
      cfma(w.d, rb[k], cb[k]);
    }
    const M2 x = rule(lq_project_to_su2_unorm(w));
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      cx na = cmk(0.0, 0.0), nb = cmk(0.0, 0.0);
      cfma(na, x.a, ra[j]);
      cfma(na, x.b, rb[j]);
      cfma(nb, x.c, ra[j]);
      cfma(nb, x.d, rb[j]);
      cur.e[j] = which == 2 ? cur.e[j] : na;
      cur.e[3 + j] = which == 0 ? nb : (which == 2 ? na : cur.e[3 + j]);
      cur.e[6 + j] = which == 0 ? cur.e[6 + j] : nb;
    }
  }
  return cur;
}
template <class Rule>
LQ_HD M3 lq_subgroup_update(const M3& u, const M3& a, Rule rule) {
  return lq_subgroup_update_acc(u, LqStapleRegs{a}, rule);
}
struct LqHeatBathRule {
  double coupling;
  LqStream* rng;
  int flags;
  LQ_HD M2 operator()(const M2& p) const { return lq_heat_bath_su2(p, coupling, *rng, flags); }
};
LQ_HD M3 lq_heat_bath_link(const M3& u, const M3& a, double coupling, LqStream& rng, int flags) {
  return lq_subgroup_update(u, a, LqHeatBathRule{coupling, &rng, flags});
}
// Option beyond the crate: over-relaxation inside the same three SU(2) sub-groups (Brown-Woch): with
// m = project(w)/k in SU(2), left-multiplying by (m^dagger)^2 reflects the block to m^dagger -- Re tr W unchanged, the
// link stays in SU(3) (the crate's SVD variants return U(3) matrices, overrelaxation.rs:96-97).
struct LqOverrelaxSu2Rule {
  LQ_HD M2 operator()(const M2& p) const {
    const double k = sqrt(m2_det(p).x);
    M2 r;
    if (!lq_is_normal(k)) {
      r.a = r.d = cmk(1.0, 0.0);
      r.b = r.c = cmk(0.0, 0.0);
      return r;
    }
    M2 v = m2_adj(p);
    v.a = cmk(v.a.x / k, v.a.y / k);
    v.b = cmk(v.b.x / k, v.b.y / k);
    v.c = cmk(v.c.x / k, v.c.y / k);
    v.d = cmk(v.d.x / k, v.d.y / k);
    return m2_mul(v, v);
  }
};
// orthonormalize_matrix (su3.rs:279-303) for the proposals: the arithmetic of lq_orthonormalize with each column scaled by
// one reciprocal square root instead of six divisions (the deterministic reprojection keeps the literal form)
LQ_HD M3 lq_orthonormalize_rs(const M3& a) {
  cx v1[3] = {a.e[0], a.e[3], a.e[6]};
  cx v2[3] = {a.e[1], a.e[4], a.e[7]};
  const double q1 = cnorm2(v1[0]) + cnorm2(v1[1]) + cnorm2(v1[2]);
  if (q1 > LQ_EPS * LQ_EPS) {
    const double s1 = LQ_RSQRT(q1);
#pragma unroll
    for (int k = 0; k < 3; ++k) v1[k] = cmk(v1[k].x * s1, v1[k].y * s1);
  }
  cx d = cmk(0.0, 0.0);
#pragma unroll
  for (int k = 0; k < 3; ++k) cfma_ca(d, v1[k], v2[k]);
  cx w[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) w[k] = csub(v2[k], cmul(v1[k], d));
  const double q2 = cnorm2(w[0]) + cnorm2(w[1]) + cnorm2(w[2]);
  if (q2 > LQ_EPS * LQ_EPS) {
    const double s2 = LQ_RSQRT(q2);
#pragma unroll
    for (int k = 0; k < 3; ++k) w[k] = cmk(w[k].x * s2, w[k].y * s2);
  }
  cx a1[3] = {cconj(v1[0]), cconj(v1[1]), cconj(v1[2])};
  cx b1[3] = {cconj(w[0]), cconj(w[1]), cconj(w[2])};
  cx cr[3] = {csub(cmul(a1[1], b1[2]), cmul(a1[2], b1[1])), csub(cmul(a1[2], b1[0]), cmul(a1[0], b1[2])),
              csub(cmul(a1[0], b1[1]), cmul(a1[1], b1[0]))};
  M3 r;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    r.e[3 * k + 0] = v1[k];
    r.e[3 * k + 1] = w[k];
    r.e[3 * k + 2] = cr[k];
  }
  return r;
}
// MetropolisHastingsSweep::potential_modif, metropolis_hastings_sweep.rs:126-143
LQ_HD M3 lq_metropolis_proposal(const M3& old_link, int n_update, double spread, LqStream& rng, int flags) {
  M3 nl = old_link;
  for (int k = 0; k < n_update; ++k) {
    M3 rm = lq_orthonormalize_rs(lq_random_su3_close_to_unity(spread, rng, flags));
    nl = m3_mul_nn(rm, nl);
  }
  return nl;
}
// delta_s_old_new_cmp, monte_carlo/mod.rs:324-334:  -Re Tr((U' - U) A) beta / CA
LQ_HD double lq_delta_s(const M3& stap, const M3& nl, const M3& ol, double beta, double CA) {
  M3 d = m3_sub(nl, ol);
  double tr = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) tr += d.e[3 * i + k].x * stap.e[3 * k + i].x - d.e[3 * i + k].y * stap.e[3 * k + i].y;
  return -tr * beta / CA;
}

// ---------------------------------------------------------------------------------------------- 3x3 SVD
// nalgebra SVD::new(a, true, true) (overrelaxation.rs:95, 167) is an un-vendored dependency; the over-relaxation
// results do not depend on the SVD convention for non-degenerate singular values, so an accurate one-sided
// Jacobi (Hestenes) serves:  a = u diag(s) v^dagger.
LQ_HD void lq_svd3(const M3& a, M3& u, M3& v) {
  M3 w = a;
  v = m3_ident();
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double app = 0.0, aqq = 0.0;
        cx apq = cmk(0.0, 0.0);
        for (int k = 0; k < 3; ++k) {
          app += cnorm2(w.e[3 * k + p]);
          aqq += cnorm2(w.e[3 * k + q]);
          apq = cadd(apq, cmul(cconj(w.e[3 * k + p]), w.e[3 * k + q]));
        }
        // the rotation's scalars from reciprocals (one division-class operation each instead of five divisions and
        // four square roots; the decomposition converges to the same factors, the iterates differ in the last bits)
        const double g2 = cnorm2(apq), pq = app * aqq;
        if (g2 <= 0.0 || g2 <= 1e-34 * pq) continue;  // |apq| <= 1e-300 (underflows to 0 when squared) or <= 1e-17 sqrt(app aqq)
        const double rg = LQ_RSQRT(g2), gg = g2 * rg;  // gg = |apq|
        off = fmax(off, gg * LQ_RSQRT(pq));
        cx ph = cmk(apq.x * rg, apq.y * rg);
        double zeta = (aqq - app) * (0.5 * rg);
        double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        double c = LQ_RSQRT(1.0 + t * t), sn = c * t;
        cx phc_sn = cscale(cconj(ph), sn), ph_sn = cscale(ph, sn);
        for (int k = 0; k < 3; ++k) {
          cx wp = w.e[3 * k + p], wq = w.e[3 * k + q];
          w.e[3 * k + p] = csub(cscale(wp, c), cmul(wq, phc_sn));
          w.e[3 * k + q] = cadd(cmul(wp, ph_sn), cscale(wq, c));
          cx vp = v.e[3 * k + p], vq = v.e[3 * k + q];
          v.e[3 * k + p] = csub(cscale(vp, c), cmul(vq, phc_sn));
          v.e[3 * k + q] = cadd(cmul(vp, ph_sn), cscale(vq, c));
        }
      }
    if (off < 1e-15) break;
  }
  u = m3_zero();
  for (int j = 0; j < 3; ++j) {
    double n2 = 0.0;
    for (int k = 0; k < 3; ++k) n2 += cnorm2(w.e[3 * k + j]);
    const double rn = n2 > 0.0 ? LQ_RSQRT(n2) : 0.0;
    for (int k = 0; k < 3; ++k)
      u.e[3 * k + j] = (n2 > 0.0) ? cmk(w.e[3 * k + j].x * rn, w.e[3 * k + j].y * rn) : cmk(k == j ? 1.0 : 0.0, 0.0);
  }
}
// su3::reverse, su3.rs:705-714: negate the off-diagonal entries
LQ_HD M3 lq_reverse(const M3& a) {
  M3 r;
#pragma unroll
  for (int k = 0; k < 9; ++k) r.e[k] = (k == 0 || k == 4 || k == 8) ? a.e[k] : cneg(a.e[k]);
  return r;
}
// OverrelaxationSweepRotation::get_modif, overrelaxation.rs:86-98:  rot U^dagger rot,  rot = u v^dagger of svd(A^dagger)
// OverrelaxationSweepReverse::get_modif, overrelaxation.rs:158-171: u reverse(u^dagger U v) v^dagger
LQ_HD M3 lq_overrelax_link(const M3& ulink, const M3& stap, int kind) {
  if (kind == 2) return lq_subgroup_update(ulink, stap, LqOverrelaxSu2Rule{});
  M3 u, v;
  lq_svd3(m3_adj(stap), u, v);
  if (kind == 0) {
    M3 rot = m3_mul_nd(u, v);
    return m3_mul_nn(m3_mul_nd(rot, ulink), rot);
  }
  M3 inner = m3_mul_nn(m3_mul_dn(u, ulink), v);
  return m3_mul_nd(m3_mul_nn(u, lq_reverse(inner)), v);
}
