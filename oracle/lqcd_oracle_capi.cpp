// lqcd_oracle_capi.cpp — extern "C" surface of the CPU ORACLE (test infrastructure only).
// See lqcd_oracle.hpp for the reference citations.  Loaded through ctypes by oracle/oracle.py.
#include "lqcd_oracle.hpp"

#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace lqo;

namespace {
inline Lattice mk(int D, const int64_t* ext, double a) { return make_lattice(D, ext, a); }
inline Dir sdir(int signed_dir) {
  // encoding: +(i+1) = positive direction i, -(i+1) = negative direction i
  return signed_dir > 0 ? Dir{signed_dir - 1, true} : Dir{-signed_dir - 1, false};
}
int levi_civita3(int a, int b, int c) {
  if (a == b || b == c || a == c) return 0;
  int p[3] = {a, b, c}, inv = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (p[i] > p[j]) ++inv;
  return inv % 2 == 0 ? 1 : -1;
}
}  // namespace

extern "C" {

int lqo_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void lqo_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

void lqo_set_flags(int flags) { g_flags = flags; }
int lqo_get_flags() { return g_flags; }

// ---- algebra ----
void lqo_generator(int a, double* out18) { store3(out18, generator(a)); }
void lqo_adjoint_to_matrix(const double* e8, double* out18) { store3(out18, adjoint_to_matrix(e8)); }
void lqo_su3_exp_i(const double* e8, double* out18) { store3(out18, su3_exp_i(e8)); }
void lqo_orthonormalize(const double* in18, double* out18) { store3(out18, orthonormalize(load3(in18))); }
void lqo_matmul(const double* a18, const double* b18, double* out18) { store3(out18, load3(a18) * load3(b18)); }
void lqo_det(const double* a18, double* out2) {
  cplx d = det(load3(a18));
  out2[0] = d.re;
  out2[1] = d.im;
}
void lqo_svd3(const double* a18, double* u18, double* s3, double* v18) {
  Mat3 u, v;
  svd3(load3(a18), u, s3, v);
  store3(u18, u);
  store3(v18, v);
}
void lqo_overrelax_link(const double* u18, const double* staple18, int kind, double* out18) {
  Mat3 u = load3(u18), a = load3(staple18);
  store3(out18, overrelax_any(u, a, kind));
}
void lqo_project_to_su2_unorm(const double* in8, double* out8) {
  // 2x2 row-major (re,im) in / out
  Mat2 m, r;
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) m.m[i][j] = {in8[2 * (i * 2 + j)], in8[2 * (i * 2 + j) + 1]};
  r = project_to_su2_unorm(m);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) {
      out8[2 * (i * 2 + j)] = r.m[i][j].re;
      out8[2 * (i * 2 + j) + 1] = r.m[i][j].im;
    }
}

// ---- rng ----
void lqo_philox_block(const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4) { Philox::block(ctr4, key2, out4); }
void lqo_stream_uniform01(uint64_t seed, uint64_t counter, uint64_t idx, int64_t n, double* out) {
  Stream st(seed, counter, idx);
  for (int64_t i = 0; i < n; ++i) out[i] = st.uniform01();
}
void lqo_random_su3_close_to_unity(uint64_t seed, uint64_t counter, uint64_t idx, double spread, double* out18) {
  Stream st(seed, counter, idx);
  store3(out18, random_su3_close_to_unity(spread, st));
}
void lqo_heat_bath_norm_samples(uint64_t seed, uint64_t counter, double param, int64_t n, double* out) {
  Stream st(seed, counter, 0);
  for (int64_t i = 0; i < n; ++i) out[i] = heat_bath_norm(param, st);
}

// ---- geometry / local quantities ----
void lqo_sij(int D, const int64_t* ext, double a, const double* U, int64_t site, int sdir_i, int sdir_j, double* out18) {
  Lattice L = mk(D, ext, a);
  store3(out18, sij(L, U, site, sdir(sdir_i), sdir(sdir_j)));
}
void lqo_pij(int D, const int64_t* ext, double a, const double* U, int64_t site, int sdir_i, int sdir_j, double* out18) {
  Lattice L = mk(D, ext, a);
  store3(out18, pij(L, U, site, sdir(sdir_i), sdir(sdir_j)));
}
void lqo_clover(int D, const int64_t* ext, double a, const double* U, int64_t site, int sdir_i, int sdir_j, double* out18) {
  Lattice L = mk(D, ext, a);
  store3(out18, clover(L, U, site, sdir(sdir_i), sdir(sdir_j)));
}
void lqo_f_mu_nu(int D, const int64_t* ext, double a, const double* U, int64_t site, int sdir_i, int sdir_j, double* out18) {
  Lattice L = mk(D, ext, a);
  store3(out18, f_mu_nu(L, U, site, sdir(sdir_i), sdir(sdir_j)));
}
// magnetic_field, field.rs:852-882 (D = 3 only: levi_civita of three indices)
void lqo_magnetic_field(int D, const int64_t* ext, double a, const double* U, int64_t site, int dir, double* out18) {
  Lattice L = mk(D, ext, a);
  Mat3 sum = zero3();
  for (int i = 0; i < D; ++i) {
    Mat3 si = zero3();
    for (int j = 0; j < D; ++j) {
      Mat3 f = f_mu_nu(L, U, site, {i, true}, {j, true});
      si = si + f * C((double)levi_civita3(dir, i, j));
    }
    sum = sum + si;
  }
  // sum / Complex(0, 2):  (x + iy) / (2i) = (y - ix) / 2
  Mat3 r;
  for (int p = 0; p < 3; ++p)
    for (int q = 0; q < 3; ++q) r.m[p][q] = {sum.m[p][q].im / 2.0, -sum.m[p][q].re / 2.0};
  store3(out18, r);
}
void lqo_staples(int D, const int64_t* ext, double a, const double* U, double* out) {
  Lattice L = mk(D, ext, a);
#pragma omp parallel for schedule(static)
  for (int64_t x = 0; x < L.ns; ++x)
    for (int d = 0; d < D; ++d) store3(out + (x * D + d) * 18, staple(L, U, x, d));
}
double lqo_delta_s(const double* staple18, const double* new18, const double* old18, double beta, double CA) {
  return delta_s(load3(staple18), load3(new18), load3(old18), beta, CA);
}

// ---- observables ----
void lqo_plaquette_sum(int D, const int64_t* ext, double a, const double* U, double* out2) {
  Lattice L = mk(D, ext, a);
  cplx s = plaquette_sum(L, U);
  out2[0] = s.re;
  out2[1] = s.im;
}
double lqo_hamiltonian_links(int D, const int64_t* ext, double a, const double* U, double beta, double CA) {
  Lattice L = mk(D, ext, a);
  return hamiltonian_links(L, U, beta, CA);
}
double lqo_hamiltonian_efield(int D, const int64_t* ext, double a, const double* E, double beta) {
  Lattice L = mk(D, ext, a);
  return hamiltonian_efield(L, E, beta);
}

// ---- molecular dynamics ----
void lqo_force(int D, const int64_t* ext, double a, const double* U, double CA, double* F, int literal) {
  Lattice L = mk(D, ext, a);
  if (literal) {
    force_field(L, U, F, CA);
  } else {
#pragma omp parallel for schedule(static)
    for (int64_t x = 0; x < L.ns; ++x)
      for (int i = 0; i < D; ++i) derivative_e_opt(L, U, x, i, CA, F + (x * D + i) * 8);
  }
}
void lqo_efield_step(int D, const int64_t* ext, double a, const double* U, double* E, double dt, double CA, int literal) {
  Lattice L = mk(D, ext, a);
  std::vector<double> out((size_t)L.nl() * 8);
  efield_step(L, U, E, out.data(), dt, CA, literal != 0);
  std::memcpy(E, out.data(), out.size() * sizeof(double));
}
void lqo_link_step(int D, const int64_t* ext, double a, double* U, const double* E, double dt, double CA) {
  Lattice L = mk(D, ext, a);
  std::vector<double> out((size_t)L.nl() * 18);
  link_step(L, U, E, out.data(), dt, CA);
  std::memcpy(U, out.data(), out.size() * sizeof(double));
}
// n repetitions of one integrator composition (kind: 0 sync_sync, 1 leap_leap, 2 sync_leap, 3 leap_sync, 4 symplectic)
void lqo_integrate(int D, const int64_t* ext, double a, double* U, double* E, int kind, double dt, int64_t n, double CA,
                   int literal) {
  Lattice L = mk(D, ext, a);
  std::vector<double> u(U, U + (size_t)L.nl() * 18), e(E, E + (size_t)L.nl() * 8);
  for (int64_t k = 0; k < n; ++k) integrate(L, u, e, kind, dt, CA, literal != 0);
  std::memcpy(U, u.data(), u.size() * sizeof(double));
  std::memcpy(E, e.data(), e.size() * sizeof(double));
}
// simulate_using_leapfrog_n, state.rs:321-358: sync_leap, (n-1) x leap_leap, leap_sync
void lqo_leapfrog_n(int D, const int64_t* ext, double a, double* U, double* E, double dt, int64_t n, double CA, int literal) {
  Lattice L = mk(D, ext, a);
  std::vector<double> u(U, U + (size_t)L.nl() * 18), e(E, E + (size_t)L.nl() * 8);
  integrate(L, u, e, SYNC_LEAP, dt, CA, literal != 0);
  for (int64_t k = 0; k + 1 < n; ++k) integrate(L, u, e, LEAP_LEAP, dt, CA, literal != 0);
  integrate(L, u, e, LEAP_SYNC, dt, CA, literal != 0);
  std::memcpy(U, u.data(), u.size() * sizeof(double));
  std::memcpy(E, e.data(), e.size() * sizeof(double));
}
// optional exponential link update  U <- exp(i dt sqrt(2 CA)/a * E^a T_a) U  (north_star wording; NOT the
// reference default, integrator/mod.rs:230-232 is Euler).  Uses su3_exp_i, su3.rs:832-855.
void lqo_link_step_exp(int D, const int64_t* ext, double a, double* U, const double* E, double dt, double CA) {
  Lattice L = mk(D, ext, a);
  double sc = dt * std::sqrt(2.0 * CA) / L.a;
#pragma omp parallel for schedule(static)
  for (int64_t l = 0; l < L.nl(); ++l) {
    double e[8];
    for (int k = 0; k < 8; ++k) e[k] = E[l * 8 + k] * sc;
    store3(U + l * 18, su3_exp_i(e) * load3(U + l * 18));
  }
}

// ---- Gauss law ----
void lqo_gauss_field(int D, const int64_t* ext, double a, const double* U, const double* E, double* out) {
  Lattice L = mk(D, ext, a);
#pragma omp parallel for schedule(static)
  for (int64_t x = 0; x < L.ns; ++x) store3(out + x * 18, gauss(L, U, E, x));
}
double lqo_gauss_sum_div(int D, const int64_t* ext, double a, const double* U, const double* E) {
  Lattice L = mk(D, ext, a);
  return gauss_sum_div(L, U, E);
}
void lqo_project_to_gauss_step(int D, const int64_t* ext, double a, const double* U, double* E) {
  Lattice L = mk(D, ext, a);
  std::vector<double> out((size_t)L.nl() * 8);
  project_to_gauss_step(L, U, E, out.data());
  std::memcpy(E, out.data(), out.size() * sizeof(double));
}
int64_t lqo_project_to_gauss(int D, const int64_t* ext, double a, const double* U, double* E, int64_t max_steps) {
  Lattice L = mk(D, ext, a);
  std::vector<double> e(E, E + (size_t)L.nl() * 8);
  int64_t it = project_to_gauss(L, U, e, max_steps);
  std::memcpy(E, e.data(), e.size() * sizeof(double));
  return it;
}

// ---- reprojection ----
void lqo_normalize_links(int D, const int64_t* ext, double a, double* U) {
  Lattice L = mk(D, ext, a);
  normalize_links(L, U);
}

// ---- start configs ----
void lqo_links_random(int D, const int64_t* ext, double a, double* U, uint64_t seed, uint64_t counter) {
  Lattice L = mk(D, ext, a);
  links_random(L, U, seed, counter);
}
void lqo_momenta_refresh(int D, const int64_t* ext, double a, double* E, uint64_t seed, uint64_t counter, double sigma) {
  Lattice L = mk(D, ext, a);
  momenta_refresh(L, E, seed, counter, sigma);
}

// ---- sweeps ----
void lqo_sweep_heatbath(int D, const int64_t* ext, double a, double* U, double beta, double coupling_scale, int order,
                        uint64_t seed, uint64_t counter, int per_link) {
  Lattice L = mk(D, ext, a);
  SweepRng rng(seed, counter, per_link != 0);
  sweep_heatbath(L, U, beta, coupling_scale, order, rng);
}
void lqo_sweep_overrelax(int D, const int64_t* ext, double a, double* U, int kind, int order) {
  Lattice L = mk(D, ext, a);
  sweep_overrelax(L, U, kind, order);
}
void lqo_sweep_metropolis(int D, const int64_t* ext, double a, double* U, double beta, double CA, int n_update,
                          double spread, int order, uint64_t seed, uint64_t counter, int per_link, int64_t* n_accept,
                          double* sum_prob) {
  Lattice L = mk(D, ext, a);
  SweepRng rng(seed, counter, per_link != 0);
  sweep_metropolis(L, U, beta, CA, n_update, spread, order, rng, n_accept, sum_prob);
}
void lqo_metropolis_single_link(int D, const int64_t* ext, double a, double* U, double beta, double CA, double spread,
                                int64_t n_hits, uint64_t seed, uint64_t counter, int64_t* n_accept, double* sum_prob) {
  Lattice L = mk(D, ext, a);
  Stream st(seed, counter, 0xFFFFFFFFFFull);
  metropolis_single_link(L, U, beta, CA, spread, n_hits, st, n_accept, sum_prob);
}

// ---- HMC trajectory ----
// HybridMonteCarloDiagnostic::next_element, hybrid_monte_carlo.rs:465-471, 573-613:
//   refresh E ~ N(0, sigma) (state.rs:1093-1108; sigma = 0.5/beta unless overridden) -> project_to_gauss ->
//   H_old -> n x integrate_symplectic -> H_new -> accept w.p. clamp(exp(H_old - H_new), 0, 1).
// If `E_inout` is non-null and use_given_e != 0 the momenta are taken from E_inout (already projected or not,
// controlled by do_project) so that GPU and oracle can run from IDENTICAL momenta.
// Returns the number of Gauss projection steps (<0 on failure).
int64_t lqo_hmc_trajectory(int D, const int64_t* ext, double a, double* U, double* E_inout, int use_given_e,
                           int do_project, double beta, double CA, double sigma, double dt, int64_t n_steps,
                           uint64_t seed, uint64_t counter, int literal, double* h_old, double* h_new, double* prob,
                           int* accepted) {
  Lattice L = mk(D, ext, a);
  std::vector<double> u(U, U + (size_t)L.nl() * 18), e((size_t)L.nl() * 8);
  if (use_given_e) std::memcpy(e.data(), E_inout, e.size() * sizeof(double));
  else momenta_refresh(L, e.data(), seed, counter, sigma);
  int64_t it = 0;
  if (do_project) {
    it = project_to_gauss(L, u.data(), e);
    if (it < 0) return it;
  }
  double h0 = hamiltonian_links(L, u.data(), beta, CA) + hamiltonian_efield(L, e.data(), beta);
  for (int64_t k = 0; k < n_steps; ++k) integrate(L, u, e, SYMPLECTIC, dt, CA, literal != 0);
  double h1 = hamiltonian_links(L, u.data(), beta, CA) + hamiltonian_efield(L, e.data(), beta);
  double p = std::fmax(std::fmin(std::exp(h0 - h1), 1.0), 0.0);
  Stream acc(seed, counter, 0xFFFFFFFFFEull);
  bool ok = acc.bernoulli(p);
  *h_old = h0;
  *h_new = h1;
  *prob = p;
  *accepted = ok ? 1 : 0;
  if (ok) std::memcpy(U, u.data(), u.size() * sizeof(double));
  if (E_inout) std::memcpy(E_inout, e.data(), e.size() * sizeof(double));
  return it;
}

}  // extern "C"
