// lq_kernels.cuh -- kernel bodies (one functor per kernel) of the pure-gauge SU(3) update path.
// Each functor names the reference loop it replaces (file:line under /root/reference/src).
// Thread mapping: whole-lattice kernels take i in [0, n_items) and decode (site, dir) themselves.
#pragma once
#include "lq_common.cuh"
#include "lq_local.cuh"

#define LQ_SITES_PER_GROUP 32  // per-link kernels: a block row = 32 consecutive sites x one direction (one warp)

// i -> (site ordinal n, dir): groups of 32 sites x D dirs; lanes of a warp share `dir` and walk consecutive sites.
template <int D>
LQ_HD bool lq_decode_link(const LqGeom& g, lq_i64 i, lq_i64& n, int& dir) {
  if (i < ((lq_i64)1 << 31)) {  // 32-bit arithmetic, see lq_site
    const unsigned blk = (unsigned)i / (unsigned)(LQ_SITES_PER_GROUP * D);
    const int r = (int)((unsigned)i - blk * (unsigned)(LQ_SITES_PER_GROUP * D));
    dir = r / LQ_SITES_PER_GROUP;
    n = (lq_i64)(blk * LQ_SITES_PER_GROUP + (unsigned)(r - dir * LQ_SITES_PER_GROUP));
    return n < g.vol;
  }
  lq_i64 blk = i / (LQ_SITES_PER_GROUP * D);
  int r = (int)(i - blk * (LQ_SITES_PER_GROUP * D));
  dir = r / LQ_SITES_PER_GROUP;
  n = blk * LQ_SITES_PER_GROUP + (r - dir * LQ_SITES_PER_GROUP);
  return n < g.vol;
}
LQ_HD lq_i64 lq_link_items(const LqGeom& g) {
  return ((g.vol + LQ_SITES_PER_GROUP - 1) / LQ_SITES_PER_GROUP) * LQ_SITES_PER_GROUP * g.D;
}

// AoS (reference) <-> M3: column-major (re, im) pairs, su3.rs:24-36 / 285-286
LQ_HD M3 lq_m3_from_aos(const double* p) {
  M3 r;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) r.e[3 * rr + c] = cmk(p[2 * (c * 3 + rr)], p[2 * (c * 3 + rr) + 1]);
  return r;
}
LQ_HD void lq_m3_to_aos(double* p, const M3& m) {
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) {
      p[2 * (c * 3 + rr)] = m.e[3 * rr + c].x;
      p[2 * (c * 3 + rr) + 1] = m.e[3 * rr + c].y;
    }
}

// ---------------------------------------------------------------------------------------------- staples
// Sum over the 2(D-1) staples around link (x, mu), already daggered the way both consumers need it:
//   A(x,mu) = sum_{nu != mu} [ U_nu(x+mu) U_mu^+(x+nu) U_nu^+(x)  +  U_nu^+(x+mu-nu) U_mu^+(x-nu) U_nu(x-nu) ]
// == `staple` of monte_carlo/mod.rs:339-362 == sum_d S_{mu,d}^+ of derivative_e (state.rs:1430-1438, with
// sij of field.rs:743-758).
template <int D>
LQ_HD M3 lq_staple_sum(const cx* LQ_RESTRICT U, const LqGeom& g, const Site<D>& x, int mu) {
  M3 acc = m3_zero();
  Site<D> xpm = lq_up<D>(g, x, mu);
#pragma unroll
  for (int nu = 0; nu < D; ++nu) {
    if (nu == mu) continue;
    {  // up:  U_nu(x+mu) U_mu^+(x+nu) U_nu^+(x)
      Site<D> xpn = lq_up<D>(g, x, nu);
      M3 a = lq_load_link(U, g, nu, lq_slot<D>(g, xpm));
      M3 b = lq_load_link(U, g, mu, lq_slot<D>(g, xpn));
      M3 t = m3_mul_nd(a, b);
      M3 c = lq_load_link(U, g, nu, lq_slot<D>(g, x));
      m3_fma_nd(acc, t, c);
    }
    {  // down:  U_nu^+(x+mu-nu) U_mu^+(x-nu) U_nu(x-nu) = (U_mu(x-nu) U_nu(x+mu-nu))^+ U_nu(x-nu)
      Site<D> xmn = lq_dn<D>(g, x, nu);
      Site<D> xpmmn = lq_dn<D>(g, xpm, nu);
      M3 a = lq_load_link(U, g, mu, lq_slot<D>(g, xmn));
      M3 b = lq_load_link(U, g, nu, lq_slot<D>(g, xpmmn));
      M3 t = m3_mul_nn(a, b);
      M3 c = lq_load_link(U, g, nu, lq_slot<D>(g, xmn));
      m3_fma_dn(acc, t, c);
    }
  }
  return acc;
}
// derivative_e for one link (state.rs:1420-1448):  F^a = -sqrt(2/CA)/a * Im Tr(T_a U_mu(x) A(x,mu))
template <int D>
LQ_HD A8 lq_force_link(const cx* LQ_RESTRICT U, const LqGeom& g, const Site<D>& x, int mu, double coef) {
  M3 a = lq_staple_sum<D>(U, g, x, mu);
  M3 u = lq_load_link(U, g, mu, lq_slot<D>(g, x));
  M3 w = m3_mul_nn(u, a);
  cx tr[8];
  lq_trace_gen(w, tr);
  A8 f;
#pragma unroll
  for (int k = 0; k < 8; ++k) f.e[k] = coef * tr[k].y;
  return f;
}

// ---------------------------------------------------------------------------------------------- marshalling
template <int D>
struct KLinksFromAos {  // LatticeStateNew::new / set_link_matrix upload (state.rs:779-815)
  LqGeom g;
  const double* aos;
  cx* U;
  LQ_HD void operator()(lq_i64 i) const {
    lq_i64 n;
    int dir;
    if (!lq_decode_link<D>(g, i, n, dir)) return;
    Site<D> st = lq_site<D>(g, n);
    lq_i64 l = lq_local_index<D>(g, st);
    lq_store_link(U, g, dir, lq_slot<D>(g, st), lq_m3_from_aos(aos + (l * D + dir) * 18));
  }
};
template <int D>
struct KLinksToAos {
  LqGeom g;
  const cx* U;
  double* aos;
  LQ_HD void operator()(lq_i64 i) const {
    lq_i64 n;
    int dir;
    if (!lq_decode_link<D>(g, i, n, dir)) return;
    Site<D> st = lq_site<D>(g, n);
    lq_i64 l = lq_local_index<D>(g, st);
    lq_m3_to_aos(aos + (l * D + dir) * 18, lq_load_link_rw(U, g, dir, lq_slot<D>(g, st)));
  }
};
template <int D>
struct KEFromAos {
  LqGeom g;
  const double* aos;
  cx* E;
  LQ_HD void operator()(lq_i64 i) const {
    lq_i64 n;
    int dir;
    if (!lq_decode_link<D>(g, i, n, dir)) return;
    Site<D> st = lq_site<D>(g, n);
    lq_i64 l = lq_local_index<D>(g, st);
    A8 a;
#pragma unroll
    for (int k = 0; k < 8; ++k) a.e[k] = aos[(l * D + dir) * 8 + k];
    lq_store_e(E, g, dir, lq_slot<D>(g, st), a);
  }
};
template <int D>
struct KEToAos {
  LqGeom g;
  const cx* E;
  double* aos;
  LQ_HD void operator()(lq_i64 i) const {
    lq_i64 n;
    int dir;
    if (!lq_decode_link<D>(g, i, n, dir)) return;
    Site<D> st = lq_site<D>(g, n);
    lq_i64 l = lq_local_index<D>(g, st);
    A8 a = lq_load_e(E, g, dir, lq_slot<D>(g, st));
#pragma unroll
    for (int k = 0; k < 8; ++k) aos[(l * D + dir) * 8 + k] = a.e[k];
  }
};
template <int D>
struct KLinksCold {  // LatticeStateDefault::new_cold, state.rs:671-679
  LqGeom g;
  cx* U;
  LQ_HD void operator()(lq_i64 i) const {
    lq_i64 n;
    int dir;
    if (!lq_decode_link<D>(g, i, n, dir)) return;
    Site<D> st = lq_site<D>(g, n);
    lq_store_link(U, g, dir, lq_slot<D>(g, st), m3_ident());
  }
};
template <int D>
struct KLinksRandom {  // LinkMatrix::new_determinist, field.rs:646-659
  LqGeom g;
  cx* U;
  uint64_t seed, counter;
  LQ_HD void operator()(lq_i64 i) const {
    lq_i64 n;
    int dir;
    if (!lq_decode_link<D>(g, i, n, dir)) return;
    Site<D> st = lq_site<D>(g, n);
    LqStream rng(seed, counter, (uint64_t)(lq_global_index<D>(g, st) * D + dir));
    lq_store_link(U, g, dir, lq_slot<D>(g, st), lq_random_su3(rng));
  }
};
template <int D>
struct KMomentaRefresh {  // EField::new_determinist with Normal(0, sigma), field.rs:1086-1099; state.rs:1097
  LqGeom g;
  cx* E;
  uint64_t seed, counter;
  double sigma;
  LQ_HD void operator()(lq_i64 i) const {
    lq_i64 n;
    int dir;
    if (!lq_decode_link<D>(g, i, n, dir)) return;
    Site<D> st = lq_site<D>(g, n);
    LqStream rng(seed, counter, (uint64_t)(lq_global_index<D>(g, st) * D + dir));
    A8 a;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      double z0, z1;
      rng.normal_pair(z0, z1);
      a.e[2 * k] = sigma * z0;
      a.e[2 * k + 1] = sigma * z1;
    }
    lq_store_e(E, g, dir, lq_slot<D>(g, st), a);
  }
};

// ---------------------------------------------------------------------------------------------- observables
// average_trace_plaquette (field.rs:775-804) and hamiltonian_links (state.rs:821-849) in one pass:
//   v[0] += Re sum_{i<j} Tr P_ij(x); v[1] += Im ...; v[2] += sum_{i<j} (1 - Re Tr P_ij(x)/CA)
// P_ij(x) = U_i(x) U_j(x+i) U_i^+(x+j) U_j^+(x)  (pij/sij, field.rs:743-771)
// One item per (site, plane): groups of 32 consecutive sites x D(D-1)/2 planes, so a warp walks 32 consecutive
// sites of one plane (coalesced) and the warps of a block share the links of their sites through L1.
template <int D>
struct KPlaquette {
  static constexpr int K = 3;
  static constexpr int NPL = D * (D - 1) / 2;
  LqGeom g;
  const cx* U;
  double CA;
  static LQ_HD lq_i64 items(const LqGeom& g) { return ((g.vol + 31) / 32) * 32 * NPL; }
  LQ_HD void operator()(lq_i64 it, double* v) const {
    lq_i64 blk = it / (32 * NPL);
    int r = (int)(it - blk * (32 * NPL));
    int pl = r >> 5;
    lq_i64 n = blk * 32 + (r & 31);
    if (n >= g.vol) return;
    // plane index -> (i < j), planes ordered (0,1), (0,2), ..., (1,2), ... as the reference's double loop
    int i = 0, j = 1, cnt = 0;
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = a + 1; b < D; ++b) {
        if (cnt == pl) {
          i = a;
          j = b;
        }
        ++cnt;
      }
    Site<D> x = lq_site<D>(g, n);
    lq_i64 p = lq_slot<D>(g, x);
    Site<D> xpi = lq_up<D>(g, x, i);
    Site<D> xpj = lq_up<D>(g, x, j);
    M3 a = m3_mul_nn(lq_load_link(U, g, i, p), lq_load_link(U, g, j, lq_slot<D>(g, xpi)));
    M3 b = m3_mul_nn(lq_load_link(U, g, j, p), lq_load_link(U, g, i, lq_slot<D>(g, xpj)));
    cx t = m3_trace_nd(a, b);
    v[0] += t.x;
    v[1] += t.y;
    v[2] += 1.0 - t.x / CA;
  }
};
// hamiltonian_efield (state.rs:1370-1385): sum_x sum_i trace_squared(E_i(x)) (field.rs:164-167); beta on the host
template <int D>
struct KEfieldEnergy {
  static constexpr int K = 1;
  LqGeom g;
  const cx* E;
  LQ_HD void operator()(lq_i64 n, double* v) const {
    Site<D> x = lq_site<D>(g, n);
    lq_i64 p = lq_slot<D>(g, x);
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      A8 a = lq_load_e(E, g, i, p);
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += a.e[k] * a.e[k];
      s += t / 2.0;
    }
    v[0] += s;
  }
};

// ---------------------------------------------------------------------------------------------- field-strength observables
// Signed directions: sd = +(d+1) / -(d+1).  LinkMatrix::matrix (field.rs:726-740): U_d(x) or, for a negative direction,
// U_d^+(x - d).
template <int D>
LQ_HD Site<D> lq_shift_signed(const LqGeom& g, const Site<D>& x, int sd) {
  return sd > 0 ? lq_up<D>(g, x, sd - 1) : lq_dn<D>(g, x, -sd - 1);
}
template <int D>
LQ_HD M3 lq_link_signed(const cx* LQ_RESTRICT U, const LqGeom& g, const Site<D>& x, int sd) {
  if (sd > 0) return lq_load_link(U, g, sd - 1, lq_slot<D>(g, x));
  return m3_adj(lq_load_link(U, g, -sd - 1, lq_slot<D>(g, lq_dn<D>(g, x, -sd - 1))));
}
// pij, field.rs:761-771:  P_ij(x) = U_i(x) S_ij^+(x),  S_ij(x) = U_j(x) U_i(x+j) U_j^+(x+i)   (sij, :743-758)
template <int D>
LQ_HD M3 lq_pij_signed(const cx* LQ_RESTRICT U, const LqGeom& g, const Site<D>& x, int si, int sj) {
  M3 u_j = lq_link_signed<D>(U, g, x, sj);
  M3 u_i_pj = lq_link_signed<D>(U, g, lq_shift_signed<D>(g, x, sj), si);
  M3 u_j_pi = lq_link_signed<D>(U, g, lq_shift_signed<D>(g, x, si), sj);
  M3 s = m3_mul_nd(m3_mul_nn(u_j, u_i_pj), u_j_pi);
  return m3_mul_nd(lq_link_signed<D>(U, g, x, si), s);
}
// clover, field.rs:807-820:  P_{i,j} + P_{j,-i} + P_{-i,-j} + P_{-j,i}
template <int D>
LQ_HD M3 lq_clover_site(const cx* LQ_RESTRICT U, const LqGeom& g, const Site<D>& x, int si, int sj) {
  M3 r = lq_pij_signed<D>(U, g, x, si, sj);
  r = m3_add(r, lq_pij_signed<D>(U, g, x, sj, -si));
  r = m3_add(r, lq_pij_signed<D>(U, g, x, -si, -sj));
  r = m3_add(r, lq_pij_signed<D>(U, g, x, -sj, si));
  return r;
}
// f_mu_nu, field.rs:825-835:  (clover_ij - clover_ji) / (8 a^2)
template <int D>
LQ_HD M3 lq_fmunu_site(const cx* LQ_RESTRICT U, const LqGeom& g, const Site<D>& x, int i, int j, double a) {
  M3 m = m3_sub(lq_clover_site<D>(U, g, x, i + 1, j + 1), lq_clover_site<D>(U, g, x, j + 1, i + 1));
  const double sc = 8.0 * a * a;
  M3 r;
#pragma unroll
  for (int k = 0; k < 9; ++k) r.e[k] = cmk(m.e[k].x / sc, m.e[k].y / sc);
  return r;
}
// levi_civita, utils.rs:306-318: product over pairs j < i of sign(index[i] - index[j])
LQ_HD int lq_levi_civita3(int a, int b, int c) {
  int sab = (b > a) - (b < a), sac = (c > a) - (c < a), sbc = (c > b) - (c < b);
  return sab * sac * sbc;
}
// mode 0: clover(si, sj);  1: f_mu_nu(i, j) (positive directions);  2: magnetic_field(dir) (field.rs:851-874)
template <int D>
struct KFieldStrength {
  LqGeom g;
  const cx* U;
  double* aos;
  int mode, p0, p1;
  double a;
  LQ_HD void operator()(lq_i64 n) const {
    Site<D> x = lq_site<D>(g, n);
    M3 r;
    if (mode == 0) {
      r = lq_clover_site<D>(U, g, x, p0, p1);
    } else if (mode == 1) {
      r = lq_fmunu_site<D>(U, g, x, p0, p1, a);
    } else {
      M3 sum = m3_zero();
      for (int i = 0; i < D; ++i) {
        M3 inner = m3_zero();
        for (int j = 0; j < D; ++j) {
          const int lc = lq_levi_civita3(p0, i, j);
          if (lc == 0) continue;  // the reference adds f_mn * 0 here
          inner = m3_add(inner, m3_scale(lq_fmunu_site<D>(U, g, x, i, j, a), (double)lc));
        }
        sum = m3_add(sum, inner);
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) r.e[k] = cmk(sum.e[k].y / 2.0, -sum.e[k].x / 2.0);  // divided by 2i
    }
    lq_m3_to_aos(aos + lq_local_index<D>(g, x) * 18, r);
  }
};

// ---------------------------------------------------------------------------------------------- molecular dynamics
template <int D>
struct KStaplesToAos {  // staple(), monte_carlo/mod.rs:339-362 (parity / debug output)
  LqGeom g;
  const cx* U;
  double* aos;
  LQ_HD void operator()(lq_i64 i) const {
    lq_i64 n;
    int dir;
    if (!lq_decode_link<D>(g, i, n, dir)) return;
    Site<D> st = lq_site<D>(g, n);
    lq_i64 l = lq_local_index<D>(g, st);
    lq_m3_to_aos(aos + (l * D + dir) * 18, lq_staple_sum<D>(U, g, st, dir));
  }
};
template <int D>
struct KForceToAos {  // derivative_e, state.rs:1420-1448
  LqGeom g;
  const cx* U;
  double* aos;
  double coef;
  LQ_HD void operator()(lq_i64 i) const {
    lq_i64 n;
    int dir;
    if (!lq_decode_link<D>(g, i, n, dir)) return;
    Site<D> st = lq_site<D>(g, n);
    lq_i64 l = lq_local_index<D>(g, st);
    A8 f = lq_force_link<D>(U, g, st, dir, coef);
#pragma unroll
    for (int k = 0; k < 8; ++k) aos[(l * D + dir) * 8 + k] = f.e[k];
  }
};
// integrate_efield over the lattice (integrator/mod.rs:240-254 under symplectic_euler_rayon.rs:88-104):
// E <- E + F dt, applied `nkick` times with the same F (nkick = 2 merges the trailing dt/2 kick of one
// symplectic step with the leading dt/2 kick of the next: same U, same F, same rounding sequence).
template <int D>
struct KEfieldStep {
  LqGeom g;
  const cx* U;
  cx* E;
  double coef, dt;
  int nkick;
  LQ_HD void operator()(lq_i64 i) const {
    lq_i64 n;
    int dir;
    if (!lq_decode_link<D>(g, i, n, dir)) return;
    Site<D> st = lq_site<D>(g, n);
    A8 f = lq_force_link<D>(U, g, st, dir, coef);
    lq_i64 p = lq_slot<D>(g, st);
    A8 e = lq_load_e(E, g, dir, p);
    for (int kk = 0; kk < nkick; ++kk) {
#pragma unroll
      for (int k = 0; k < 8; ++k) e.e[k] = fma(f.e[k], dt, e.e[k]);
    }
    lq_store_e(E, g, dir, p, e);
  }
};
// integrate_link over the lattice (integrator/mod.rs:216-233 with derivative_u state.rs:1407-1417):
//   Euler:  U <- U + dt * (E.to_matrix() U) * i sqrt(2 CA) / a        (the reference's rule)
//   exp  :  U <- exp(i dt sqrt(2 CA)/a E^a T_a) U                     (optional, su3_exp_i su3.rs:832-855)
template <int D>
LQ_HD M3 lq_link_update(const M3& u, const A8& e, double dt, double c_u, int use_exp) {
  if (use_exp) {
    A8 s;
#pragma unroll
    for (int k = 0; k < 8; ++k) s.e[k] = e.e[k] * (dt * c_u);
    return m3_mul_nn(lq_su3_exp_i(s), u);
  }
  M3 eu = m3_mul_nn(lq_adjoint_to_matrix(e), u);
  M3 r;
  // (x + iy) * (i c) = -c y + i c x ; then * dt and added to U
  double f = c_u * dt;
#pragma unroll
  for (int k = 0; k < 9; ++k) r.e[k] = cmk(fma(-f, eu.e[k].y, u.e[k].x), fma(f, eu.e[k].x, u.e[k].y));
  return r;
}
template <int D>
struct KLinkStep {
  LqGeom g;
  const cx* Uin;
  cx* Uout;
  const cx* E;
  double dt, c_u;  // c_u = sqrt(2 CA) / a
  int use_exp;
  LQ_HD void operator()(lq_i64 i) const {
    lq_i64 n;
    int dir;
    if (!lq_decode_link<D>(g, i, n, dir)) return;
    Site<D> st = lq_site<D>(g, n);
    lq_i64 p = lq_slot<D>(g, st);
    M3 u = lq_load_link_rw(Uin, g, dir, p);
    A8 e = lq_load_e(E, g, dir, p);
    lq_store_link(Uout, g, dir, p, lq_link_update<D>(u, e, dt, c_u, use_exp));
  }
};
// Fused symplectic-Euler kernel: E(x,mu) += nkick * dt_e * F[U](x,mu);  Unew(x,mu) = step(U(x,mu), E_new, dt_u).
// One pass reads U once (plus cached neighbours) and E once, writes E and the second link buffer.
// Equivalent to KEfieldStep followed by KLinkStep (symplectic_euler_rayon.rs:222-252, first two stages).
template <int D>
struct KEfieldLinkStep {
  LqGeom g;
  const cx* U;
  cx* Unew;
  cx* E;
  double coef, dt_e, dt_u, c_u;
  int nkick, use_exp;
  LQ_HD void operator()(lq_i64 i) const {
    lq_i64 n;
    int dir;
    if (!lq_decode_link<D>(g, i, n, dir)) return;
    Site<D> st = lq_site<D>(g, n);
    lq_i64 p = lq_slot<D>(g, st);
    M3 a = lq_staple_sum<D>(U, g, st, dir);
    M3 u = lq_load_link(U, g, dir, p);
    M3 w = m3_mul_nn(u, a);
    cx tr[8];
    lq_trace_gen(w, tr);
    A8 e = lq_load_e(E, g, dir, p);
    for (int kk = 0; kk < nkick; ++kk) {
#pragma unroll
      for (int k = 0; k < 8; ++k) e.e[k] = fma(coef * tr[k].y, dt_e, e.e[k]);
    }
    lq_store_e(E, g, dir, p, e);
    lq_store_link(Unew, g, dir, p, lq_link_update<D>(u, e, dt_u, c_u, use_exp));
  }
};
template <int D>
struct KReunitarize {  // LinkMatrix::normalize, field.rs:897-901 -> orthonormalize_matrix su3.rs:279-303
  LqGeom g;
  cx* U;
  LQ_HD void operator()(lq_i64 i) const {
    lq_i64 n;
    int dir;
    if (!lq_decode_link<D>(g, i, n, dir)) return;
    Site<D> st = lq_site<D>(g, n);
    lq_i64 p = lq_slot<D>(g, st);
    lq_store_link(U, g, dir, p, lq_orthonormalize(lq_load_link_rw(U, g, dir, p)));
  }
};

// Peer table of the fused "compute + halo push" kernels (decomposed contexts with the peer-to-peer transport):
// threads that own an element of a boundary slice also store it straight into the ghost layer of the neighbour
// rank(s) over NVLink.  The stores are posted, so the transfer overlaps the arithmetic of the other blocks.
struct LqPush {
  cx* peer[8];      // neighbour k's buffer (the allocation that corresponds to the one being written)
  int delta[8];     // slot shift into its ghost layer
  int nbmap[3][3];  // [o_a + 1][o_b + 1] -> neighbour index or -1; o = neighbour offset in directions D-2, D-1
};
// which boundary (if any) of the two splittable directions a storage site sits on: 0 low, 1 none, 2 high
template <int D>
LQ_HD void lq_boundary_code(const LqGeom& g, const Site<D>& x, int& oa, int& ob) {
  oa = 1;
  ob = 1;
  if (D >= 3 && g.ghost[D - 2]) oa = x.x[D - 2] == 1 ? 0 : (x.x[D - 2] == g.ext[D - 2] ? 2 : 1);
  if (g.ghost[D - 1]) ob = x.x[D - 1] == 1 ? 0 : (x.x[D - 1] == g.ext[D - 1] ? 2 : 1);
}

// ---------------------------------------------------------------------------------------------- Gauss law
// EField::gauss, field.rs:1174-1195:  G(x) = sum_i [ E_i(x) - U_i^+(x-i) E_i(x-i) U_i(x-i) ]
template <int D>
LQ_HD M3 lq_gauss_site(const cx* LQ_RESTRICT U, const cx* E, const LqGeom& g, const Site<D>& x) {
  M3 acc = m3_zero();
  lq_i64 p = lq_slot<D>(g, x);
#pragma unroll
  for (int i = 0; i < D; ++i) {
    acc = m3_add(acc, lq_adjoint_to_matrix(lq_load_e(E, g, i, p)));
    Site<D> xm = lq_dn<D>(g, x, i);
    lq_i64 pm = lq_slot<D>(g, xm);
    M3 u = lq_load_link(U, g, i, pm);
    M3 em = lq_adjoint_to_matrix(lq_load_e(E, g, i, pm));
    M3 t = m3_mul_dn(u, em);  // U^+ E
    M3 neg = m3_zero();
    m3_fma_nn(neg, t, u);
    acc = m3_sub(acc, neg);
  }
  return acc;
}
// Writes G on every storage site whose backward neighbours are available: interior sites (needs the low E/U
// ghosts) -- the high ghost layer of G needed by the projection step is exchanged / recomputed by the caller.
template <int D>
struct KGaussField {
  LqGeom g;
  const cx* U;
  const cx* E;
  cx* G;
  const LqPush* ps;  // non-null: also store boundary sites into the neighbours' ghost layers (fused halo push)
  LQ_HD void operator()(lq_i64 n) const {
    Site<D> x = lq_site<D>(g, n);
    M3 m = lq_gauss_site<D>(U, E, g, x);
    lq_i64 p = lq_slot<D>(g, x);
    lq_store_g(G, p, m);
    if (ps) {
      int oa, ob;
      lq_boundary_code<D>(g, x, oa, ob);
      if (oa != 1 || ob != 1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {  // face a, face b, corner
          const int a = c == 1 ? 1 : oa, b = c == 0 ? 1 : ob;
          if ((a == 1 && b == 1) || (c == 2 && (oa == 1 || ob == 1))) continue;
          const int k = ps->nbmap[a][b];
          if (k < 0) continue;
          lq_store_g(ps->peer[k], p + ps->delta[k], m);
        }
      }
    }
  }
};
template <int D>
struct KGaussToAos {
  LqGeom g;
  const cx* G;
  double* aos;
  LQ_HD void operator()(lq_i64 n) const {
    Site<D> x = lq_site<D>(g, n);
    lq_i64 p = lq_slot<D>(g, x);
    M3 m = lq_load_g(G, p);
    lq_m3_to_aos(aos + lq_local_index<D>(g, x) * 18, m);
  }
};
// gauss_sum_div, field.rs:1199-1220:  sum_x | Tr( (sum_a T_a) G(x) ) |
template <int D>
struct KGaussDiv {
  static constexpr int K = 1;
  LqGeom g;
  const cx* G;
  LQ_HD void operator()(lq_i64 n, double* v) const {
    Site<D> x = lq_site<D>(g, n);
    lq_i64 p = lq_slot<D>(g, x);
    M3 m = lq_load_g(G, p);
    cx tr[8];
    lq_trace_gen(m, tr);
    double re = 0.0, im = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      re += tr[k].x;
      im += tr[k].y;
    }
    v[0] += sqrt(re * re + im * im);
  }
};
// project_to_gauss_step, field.rs:1301-1337:
//   E_i^a(x) <- 2 Re Tr( T_a [ (U_i(x) G(x) U_i^+(x) G(x+i) - G(x)) 0.12 + T_a E_i^a(x) ] )
//             = 2 Re Tr(T_a M) + E_i^a(x)            (Tr T_a T_a = 1/2)
LQ_HD A8 lq_gauss_project_link(const M3& u, const M3& gx, const M3& gp, A8 e) {
  M3 t = m3_mul_nd(m3_mul_nn(u, gx), u);
  M3 m = m3_mul_nn(t, gp);
  m = m3_scale(m3_sub(m, gx), 0.12);
  cx tr[8];
  lq_trace_gen(m, tr);
#pragma unroll
  for (int k = 0; k < 8; ++k) e.e[k] = 2.0 * (tr[k].x + 0.5 * e.e[k]);
  return e;
}
template <int D>
struct KGaussProjectStep {
  LqGeom g;
  const cx* U;
  const cx* G;
  const cx* Ein;
  cx* Eout;
  LQ_HD void operator()(lq_i64 i) const {
    lq_i64 n;
    int dir;
    if (!lq_decode_link<D>(g, i, n, dir)) return;
    Site<D> x = lq_site<D>(g, n);
    lq_i64 p = lq_slot<D>(g, x);
    Site<D> xp = lq_up<D>(g, x, dir);
    lq_i64 pp = lq_slot<D>(g, xp);
    M3 u = lq_load_link(U, g, dir, p);
    M3 gx = lq_load_g(G, p), gp = lq_load_g(G, pp);
    A8 e = lq_gauss_project_link(u, gx, gp, lq_load_e(Ein, g, dir, p));
    lq_store_e(Eout, g, dir, p, e);
  }
};
// The same projection step for the links (y, hd) of the LOW ghost layer of a decomposed direction hd (the only E
// entries of a ghost layer any kernel reads: EField::gauss needs E_i(x - i)).  Recomputing them here -- same
// arithmetic and inputs as the rank that owns them -- keeps them valid without an E-field exchange per iteration.
template <int D>
struct KGaussProjectGhost {
  LqGeom g;
  const cx* U;
  const cx* G;
  const cx* Ein;
  cx* Eout;
  int hd;
  LQ_HD void operator()(lq_i64 n) const {
    Site<D> y;
    lq_i64 q = n / g.ext[0];
    int lane = (int)(n - q * g.ext[0]);
    y.x[0] = lane < g.ne0 ? 2 * lane : 2 * (lane - g.ne0) + 1;
    y.s = y.x[0];
#pragma unroll
    for (int d = 1; d < D; ++d) {
      int xd = 0;
      if (d != hd) {
        lq_i64 r = q / g.ext[d];
        xd = (int)(q - r * g.ext[d]) + g.ghost[d];
        q = r;
      }
      y.x[d] = xd;
      y.s += (lq_i64)xd * g.sstride[d];
    }
    lq_i64 p = lq_slot<D>(g, y);
    lq_i64 pp = lq_slot<D>(g, lq_up<D>(g, y, hd));
    M3 u = lq_load_link(U, g, hd, p);
    M3 gy = lq_load_g(G, p), gp = lq_load_g(G, pp);
    lq_store_e(Eout, g, hd, p, lq_gauss_project_link(u, gy, gp, lq_load_e(Ein, g, hd, p)));
  }
};
// One whole iteration of project_to_gauss (field.rs:1265-1294) in ONE pass: the projection step of every link of
// site x, and the Gauss field of the projected E at x,
//   G'(x) = sum_i [ E'_i(x) - U_i^+(x-i) E'_i(x-i) U_i(x-i) ]            (field.rs:1174-1195)
// where the backward neighbours' E'_i(x-i) are RECOMPUTED here (same arithmetic, same inputs => same bits as the
// thread that owns them) instead of being read back in a second pass.  Against KGaussProjectStep + KGaussField this
// reads U once instead of twice per iteration (1376 instead of 2208 B/site).  In a decomposed direction the
// recomputed E'_i(x-i) of a low-ghost site is also stored, which keeps the only E ghost entries any kernel reads
// (component i of the low ghost in direction i) valid without an exchange.
template <int D>
struct KGaussIter {
  LqGeom g;
  const cx* U;
  const cx* Gin;
  const cx* Ein;
  cx* Eout;
  cx* Gout;
  LQ_HD void operator()(lq_i64 n) const {
    Site<D> x = lq_site<D>(g, n);
    lq_i64 p = lq_slot<D>(g, x);
    const M3 gx = lq_load_g(Gin, p);
    M3 acc = m3_zero();
#pragma unroll 1
    for (int i = 0; i < D; ++i) {
      {
        Site<D> xp = lq_up<D>(g, x, i);
        M3 u = lq_load_link(U, g, i, p);
        M3 gp = lq_load_g(Gin, lq_slot<D>(g, xp));
        A8 e = lq_gauss_project_link(u, gx, gp, lq_load_e(Ein, g, i, p));
        lq_store_e(Eout, g, i, p, e);
        acc = m3_add(acc, lq_adjoint_to_matrix(e));
      }
      {
        Site<D> xm = lq_dn<D>(g, x, i);
        lq_i64 pm = lq_slot<D>(g, xm);
        M3 um = lq_load_link(U, g, i, pm);
        M3 gm = lq_load_g(Gin, pm);
        A8 em = lq_gauss_project_link(um, gm, gx, lq_load_e(Ein, g, i, pm));
        if (g.ghost[i] && x.x[i] == 1) lq_store_e(Eout, g, i, pm, em);
        M3 t = m3_mul_dn(um, lq_adjoint_to_matrix(em));  // U^+ E
        M3 neg = m3_zero();
        m3_fma_nn(neg, t, um);
        acc = m3_sub(acc, neg);
      }
    }
    lq_store_g(Gout, p, acc);
  }
};

// ---------------------------------------------------------------------------------------------- checkerboard sweeps
// One launch per (dir, colour): the links (x, dir) with colour(x) = parity do not enter each other's staples,
// so the parallel update equals the serial loop over them (heat_bath.rs:113-123 visits links one by one).
template <int D>
struct KHeatBath {  // HeatBathSweep, heat_bath.rs:73-123
  LqGeom g;
  cx* U;
  int dir, parity, flags;
  double coupling;  // beta * coupling_scale
  uint64_t seed, counter;
  int odd_mask, cmask;  // colour classes of lattices with odd extents (lq_site_class); 0, 0 otherwise
  LQ_HD void operator()(lq_i64 n) const {
    Site<D> x;
    if (!lq_site_class<D>(g, n, parity, odd_mask, cmask, x)) return;
    lq_i64 p = lq_slot<D>(g, x);
    M3 a = lq_staple_sum<D>(U, g, x, dir);
    M3 u = lq_load_link_rw(U, g, dir, p);
    LqStream rng(seed, counter, (uint64_t)(lq_global_index<D>(g, x) * D + dir));
    lq_store_link(U, g, dir, p, lq_heat_bath_link(u, a, coupling, rng, flags));
  }
};
template <int D>
struct KOverrelax {  // OverrelaxationSweep{Rotation,Reverse}, overrelaxation.rs:86-110, 158-184
  LqGeom g;
  cx* U;
  int dir, parity, kind;
  int odd_mask, cmask;
  LQ_HD void operator()(lq_i64 n) const {
    Site<D> x;
    if (!lq_site_class<D>(g, n, parity, odd_mask, cmask, x)) return;
    lq_i64 p = lq_slot<D>(g, x);
    M3 a = lq_staple_sum<D>(U, g, x, dir);
    M3 u = lq_load_link_rw(U, g, dir, p);
    lq_store_link(U, g, dir, p, lq_overrelax_link(u, a, kind));
  }
};
template <int D>
struct KMetropolis {  // MetropolisHastingsSweep, metropolis_hastings_sweep.rs:126-174; v = (#accepted, sum prob)
  static constexpr int K = 2;
  LqGeom g;
  cx* U;
  int dir, parity, flags, n_update;
  double beta, CA, spread;
  uint64_t seed, counter;
  int odd_mask, cmask;
  LQ_HD void operator()(lq_i64 n, double* v) const {
    Site<D> x;
    if (!lq_site_class<D>(g, n, parity, odd_mask, cmask, x)) return;
    lq_i64 p = lq_slot<D>(g, x);
    M3 old = lq_load_link_rw(U, g, dir, p);
    LqStream rng(seed, counter, (uint64_t)(lq_global_index<D>(g, x) * D + dir));
    M3 prop = lq_metropolis_proposal(old, n_update, spread, rng, flags);
    M3 a = lq_staple_sum<D>(U, g, x, dir);
    double proba = fmax(fmin(exp(-lq_delta_s(a, prop, old, beta, CA)), 1.0), 0.0);
    v[1] += proba;
    if (rng.bernoulli(proba)) {
      v[0] += 1.0;
      lq_store_link(U, g, dir, p, prop);
    }
  }
};

// Random single-link Metropolis hits (MetropolisHastingsDeltaDiagnostic::next_element, metropolis_hastings.rs:374-417:
// ONE uniformly random link per call; proposal orthonormalize(random_su3_close_to_unity(spread)) * U, accepted with
// probability min(1, exp(-dS))), batched: the n_hits hits of one call all sit on links of one (direction, colour)
// class drawn by the host, so no two of them enter each other's staples, and each hit draws its site of that colour
// from its own Philox stream (seed, counter, hit index).  Two hits on the same link: the lower hit index keeps it
// (phase 0 claims with an atomic minimum, phase 1 performs the claimed hits), the others are dropped and reported.
// With n_hits = 1 this is the reference's call: a uniformly random link.
#if defined(LQ_HOST_EMU)
#define LQ_ATOMIC_MIN(ptr, val)                \
  do {                                         \
    _Pragma("omp critical(lq_claim)") {        \
      if ((val) < *(ptr)) *(ptr) = (val);      \
    }                                          \
  } while (0)
#elif defined(__CUDA_ARCH__)
#define LQ_ATOMIC_MIN(ptr, val) atomicMin((ptr), (val))
#else
#define LQ_ATOMIC_MIN(ptr, val) ((void)0)
#endif
template <int D>
struct KMetropolisHits {
  static constexpr int K = 3;  // (#accepted, sum of acceptance probabilities, #performed)
  LqGeom g;
  cx* U;
  int* claim;  // one entry per site of the colour (vol / 2), preset to INT_MAX
  int phase, dir, parity, flags, force_accept;
  double beta, CA, spread;
  uint64_t seed, counter;
  lq_i64 hit0;  // index of the first hit of this launch (sequential mode launches one hit at a time)
  LQ_HD void operator()(lq_i64 i, double* v) const {
    LqStream rng(seed, counter, (uint64_t)(hit0 + i));
    Site<D> x;
    int dir = this->dir;
    if (claim == nullptr) {
      // sequential mode (lattices with an odd extent have no two-colour classes): one hit per launch on a uniformly
      // random link -- the reference's own sequence of calls
      lq_i64 n = (lq_i64)(rng.uniform01() * (double)g.vol);
      if (n >= g.vol) n = g.vol - 1;
      dir = (int)(rng.uniform01() * D);
      if (dir >= D) dir = D - 1;
      x = lq_site<D>(g, n);
    } else {
      const lq_i64 half = g.vol >> 1;
      lq_i64 n = (lq_i64)(rng.uniform01() * (double)half);
      if (n >= half) n = half - 1;
      if (phase == 0) {
        LQ_ATOMIC_MIN(claim + n, (int)i);
        return;
      }
      if (claim[n] != (int)i) return;
      x = lq_site_eo<D>(g, n, parity);
    }
    v[2] += 1.0;
    const lq_i64 p = lq_slot<D>(g, x);
    const M3 old = lq_load_link_rw(U, g, dir, p);
    const M3 prop = lq_metropolis_proposal(old, 1, spread, rng, flags);
    if (force_accept) {  // MetropolisHastings::potential_next_element: the caller accepts on the Hamiltonians
      v[0] += 1.0;
      v[1] += 1.0;
      lq_store_link(U, g, dir, p, prop);
      return;
    }
    const M3 a = lq_staple_sum<D>(U, g, x, dir);
    const double proba = fmax(fmin(exp(-lq_delta_s(a, prop, old, beta, CA)), 1.0), 0.0);
    v[1] += proba;
    if (rng.bernoulli(proba)) {
      v[0] += 1.0;
      lq_store_link(U, g, dir, p, prop);
    }
  }
};
template <class F>
struct KNoReduce {  // runs a reduction functor for its side effects only
  F f;
  LQ_HD void operator()(lq_i64 i) const {
    double v[F::K];
#pragma unroll
    for (int k = 0; k < F::K; ++k) v[k] = 0.0;
    f(i, v);
  }
};
struct KFillInt {
  int* p;
  int val;
  LQ_HD void operator()(lq_i64 i) const { p[i] = val; }
};

// ---------------------------------------------------------------------------------------------- halos
// Face slices of a decomposed direction `hd`: n in [0, face sites) enumerates storage sites with x[hd] fixed
// (all other directions over their full STORAGE extent, so later directions carry earlier ghosts = corners).
template <int D>
LQ_HD lq_i64 lq_face_site(const LqGeom& g, int hd, int xh, lq_i64 n) {
  lq_i64 s = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    int xd;
    if (d == hd) {
      xd = xh;
    } else {
      lq_i64 q = n / g.sext[d];
      xd = (int)(n - q * g.sext[d]);
      n = q;
    }
    s += (lq_i64)xd * g.sstride[d];
  }
  return s;
}
// planes = 9*D (links) or 4*D (efield); buffer layout [plane][face site]
template <int D>
struct KHaloPack {
  LqGeom g;
  const cx* F;
  cx* buf;
  int hd, xh, planes;
  lq_i64 nface;
  LQ_HD void operator()(lq_i64 i) const {
    lq_i64 pl = i / nface;
    lq_i64 n = i - pl * nface;
    if (pl >= planes) return;
    buf[i] = F[lq_addr(lq_slot_s(g, lq_face_site<D>(g, hd, xh, n)), planes, (int)pl)];
  }
};
template <int D>
struct KHaloUnpack {
  LqGeom g;
  cx* F;
  const cx* buf;
  int hd, xh, planes;
  lq_i64 nface;
  LQ_HD void operator()(lq_i64 i) const {
    lq_i64 pl = i / nface;
    lq_i64 n = i - pl * nface;
    if (pl >= planes) return;
    F[lq_addr(lq_slot_s(g, lq_face_site<D>(g, hd, xh, n)), planes, (int)pl)] = buf[i];
  }
};
