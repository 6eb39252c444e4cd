// lqcd_oracle.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A plain C++17 restatement of the pure-gauge SU(3) update hot path of
// lattice-qcd-rs v0.2.1 (reference mounted at /root/reference, Rust, cannot be
// compiled in this image: no rustc/cargo).  Every function cites the reference
// file:line it follows.  Layouts are the reference's AoS layouts:
//   links : Nl * 18 f64, link index = site*D + dir, matrix column-major
//           (nalgebra ArrayStorage), complex = (re, im)          field.rs:584-586
//   efield: Nl * 8 f64, flat index (site*D+dir)*8 + a             field.rs:1025-1027
//   site  : sum_k x_k * stride_k, x_0 fastest                     lattice.rs:909-916
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may call this.  The product path (lattice_qcd_rs_b200/csrc)
// never includes or links it.
//
// Parity pin status: the reference holds no golden numbers for RNG streams
// (rand/rand_distr are un-vendored: "parity unpinned" for the stochastic
// streams); deterministic paths are pinned by the reference's own known-answer
// tests, restated in tests/test_oracle_golden.py.
//
// Build: g++ -O3 -march=native -ffp-contract=off -fopenmp (Rust never contracts
// a*b+c into an FMA, so neither may the oracle).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace lqo {

constexpr int MAXD = 8;
constexpr double EPS = std::numeric_limits<double>::epsilon();
constexpr double PI = 3.14159265358979323846264338327950288;

// ---------------------------------------------------------------- complex
// num-complex semantics: (a+bi)(c+di) = (ac-bd) + (ad+bc)i, no inf/nan fixups.
struct cplx {
  double re, im;
};
inline cplx C(double re, double im = 0.0) { return {re, im}; }
inline cplx operator+(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
inline cplx operator-(cplx a, cplx b) { return {a.re - b.re, a.im - b.im}; }
inline cplx operator-(cplx a) { return {-a.re, -a.im}; }
inline cplx operator*(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
inline cplx operator*(cplx a, double s) { return {a.re * s, a.im * s}; }
inline cplx operator/(cplx a, double s) { return {a.re / s, a.im / s}; }
inline cplx conj(cplx a) { return {a.re, -a.im}; }
inline double norm2(cplx a) { return a.re * a.re + a.im * a.im; }
inline double cabs(cplx a) { return std::hypot(a.re, a.im); }

// ---------------------------------------------------------------- 3x3 / 2x2
struct Mat3 {
  cplx m[3][3];  // m[row][col]
};
struct Mat2 {
  cplx m[2][2];
};

inline Mat3 zero3() {
  Mat3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = C(0);
  return r;
}
inline Mat3 ident3() {
  Mat3 r = zero3();
  for (int i = 0; i < 3; ++i) r.m[i][i] = C(1);
  return r;
}
inline Mat3 operator+(const Mat3& a, const Mat3& b) {
  Mat3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] + b.m[i][j];
  return r;
}
inline Mat3 operator-(const Mat3& a, const Mat3& b) {
  Mat3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] - b.m[i][j];
  return r;
}
// nalgebra gemm for small complex matrices: c_ij = sum_k a_ik b_kj, k ascending.
inline Mat3 operator*(const Mat3& a, const Mat3& b) {
  Mat3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      cplx s = a.m[i][0] * b.m[0][j];
      s = s + a.m[i][1] * b.m[1][j];
      s = s + a.m[i][2] * b.m[2][j];
      r.m[i][j] = s;
    }
  return r;
}
inline Mat3 operator*(const Mat3& a, cplx s) {
  Mat3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] * s;
  return r;
}
inline Mat3 adj(const Mat3& a) {
  Mat3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = conj(a.m[j][i]);
  return r;
}
inline cplx trace(const Mat3& a) { return a.m[0][0] + a.m[1][1] + a.m[2][2]; }
inline cplx det(const Mat3& a) {
  // cofactor expansion along the first row (nalgebra Matrix3::determinant)
  cplx m11 = a.m[0][0], m12 = a.m[0][1], m13 = a.m[0][2];
  cplx m21 = a.m[1][0], m22 = a.m[1][1], m23 = a.m[1][2];
  cplx m31 = a.m[2][0], m32 = a.m[2][1], m33 = a.m[2][2];
  cplx minor1 = m22 * m33 - m32 * m23;
  cplx minor2 = m21 * m33 - m31 * m23;
  cplx minor3 = m21 * m32 - m31 * m22;
  return m11 * minor1 - m12 * minor2 + m13 * minor3;
}
inline double frob(const Mat3& a) {
  double s = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) s += norm2(a.m[i][j]);
  return std::sqrt(s);
}

inline Mat2 operator*(const Mat2& a, const Mat2& b) {
  Mat2 r;
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j];
  return r;
}
inline Mat2 adj(const Mat2& a) {
  Mat2 r;
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) r.m[i][j] = conj(a.m[j][i]);
  return r;
}
inline cplx det(const Mat2& a) { return a.m[0][0] * a.m[1][1] - a.m[1][0] * a.m[0][1]; }
inline cplx trace(const Mat2& a) { return a.m[0][0] + a.m[1][1]; }

// AoS load / store: column-major, (re, im) pairs.            su3.rs:24-36, 285-286
inline Mat3 load3(const double* p) {
  Mat3 r;
  for (int c = 0; c < 3; ++c)
    for (int rr = 0; rr < 3; ++rr) r.m[rr][c] = {p[2 * (c * 3 + rr)], p[2 * (c * 3 + rr) + 1]};
  return r;
}
inline void store3(double* p, const Mat3& a) {
  for (int c = 0; c < 3; ++c)
    for (int rr = 0; rr < 3; ++rr) {
      p[2 * (c * 3 + rr)] = a.m[rr][c].re;
      p[2 * (c * 3 + rr) + 1] = a.m[rr][c].im;
    }
}

// ---------------------------------------------------------------- generators
// Gell-Mann / 2, su3.rs:24-194 (order GENERATOR_1..8).
inline Mat3 generator(int a) {
  constexpr double S3 = 0.288675134594812900;   // ONE_OVER_2_SQRT_3        su3.rs:183
  constexpr double M3 = -0.577350269189625800;  // MINUS_ONE_OVER_SQRT_3    su3.rs:182
  Mat3 g = zero3();
  switch (a) {
    case 0: g.m[0][1] = C(0.5); g.m[1][0] = C(0.5); break;
    case 1: g.m[0][1] = C(0, -0.5); g.m[1][0] = C(0, 0.5); break;
    case 2: g.m[0][0] = C(0.5); g.m[1][1] = C(-0.5); break;
    case 3: g.m[0][2] = C(0.5); g.m[2][0] = C(0.5); break;
    case 4: g.m[0][2] = C(0, -0.5); g.m[2][0] = C(0, 0.5); break;
    case 5: g.m[1][2] = C(0.5); g.m[2][1] = C(0.5); break;
    case 6: g.m[1][2] = C(0, -0.5); g.m[2][1] = C(0, 0.5); break;
    case 7: g.m[0][0] = C(S3); g.m[1][1] = C(S3); g.m[2][2] = C(M3); break;
  }
  return g;
}

// Su3Adjoint::to_matrix, field.rs:106-112:  sum_a GENERATORS[a] * Complex(e_a)
inline Mat3 adjoint_to_matrix(const double* e) {
  Mat3 s = zero3();
  for (int a = 0; a < 8; ++a) s = s + generator(a) * C(e[a]);
  return s;
}
// Su3Adjoint::trace_squared, field.rs:164-167
inline double trace_squared(const double* e) {
  double s = 0;
  for (int a = 0; a < 8; ++a) s += e[a] * e[a];
  return s / 2.0;
}

// ---------------------------------------------------------------- geometry
// LatticeCyclic generalised to per-direction extents (the reference has one
// `dim` for all directions, lattice.rs:44-49).
struct Lattice {
  int D = 4;
  int64_t ext[MAXD];
  int64_t stride[MAXD];
  int64_t ns = 0;  // number_of_points
  double a = 1.0;  // lattice spacing `size`
  int64_t nl() const { return ns * D; }
};
inline Lattice make_lattice(int D, const int64_t* ext, double a) {
  Lattice L;
  L.D = D;
  L.a = a;
  int64_t s = 1;
  for (int k = 0; k < D; ++k) {
    L.ext[k] = ext[k];
    L.stride[k] = s;
    s *= ext[k];
  }
  L.ns = s;
  return L;
}
struct Dir {
  int idx;
  bool pos;
};
inline Dir operator-(Dir d) { return {d.idx, !d.pos}; }
// add_point_direction, lattice.rs:303-323 (shift by one with periodic wrap)
inline int64_t shift(const Lattice& L, int64_t site, Dir d) {
  int64_t x = (site / L.stride[d.idx]) % L.ext[d.idx];
  int64_t xn = d.pos ? (x + 1) % L.ext[d.idx] : (x == 0 ? L.ext[d.idx] - 1 : x - 1);
  return site + (xn - x) * L.stride[d.idx];
}
// LinkMatrix::matrix, field.rs:726-740 with link_canonical lattice.rs:84-97:
//   U_{-i}(x) = U_i(x - i)^dagger
inline Mat3 link(const Lattice& L, const double* U, int64_t site, Dir d) {
  if (d.pos) return load3(U + (site * L.D + d.idx) * 18);
  int64_t s = shift(L, site, d);
  return adj(load3(U + (s * L.D + d.idx) * 18));
}
// sij, field.rs:743-758: S_ij(x) = U_j(x) U_i(x+j) U_j^dagger(x+i)
inline Mat3 sij(const Lattice& L, const double* U, int64_t x, Dir i, Dir j) {
  Mat3 u_j = link(L, U, x, j);
  Mat3 u_i_pj = link(L, U, shift(L, x, j), i);
  Mat3 u_j_pi_d = adj(link(L, U, shift(L, x, i), j));
  return u_j * u_i_pj * u_j_pi_d;
}
// pij, field.rs:761-771: P_ij(x) = U_i(x) S_ij^dagger(x)
inline Mat3 pij(const Lattice& L, const double* U, int64_t x, Dir i, Dir j) {
  Mat3 s = sij(L, U, x, i, j);
  Mat3 u_i = link(L, U, x, i);
  return u_i * adj(s);
}
// clover, field.rs:807-820
inline Mat3 clover(const Lattice& L, const double* U, int64_t x, Dir i, Dir j) {
  return pij(L, U, x, i, j) + pij(L, U, x, j, -i) + pij(L, U, x, -i, -j) + pij(L, U, x, -j, i);
}
// f_mu_nu, field.rs:825-835
inline Mat3 f_mu_nu(const Lattice& L, const double* U, int64_t x, Dir i, Dir j) {
  Mat3 m = clover(L, U, x, i, j) - clover(L, U, x, j, i);
  double s = 8.0 * L.a * L.a;
  Mat3 r;
  for (int p = 0; p < 3; ++p)
    for (int q = 0; q < 3; ++q) r.m[p][q] = m.m[p][q] / s;
  return r;
}

// ---------------------------------------------------------------- observables
// average_trace_plaquette numerator, field.rs:775-804: sum_x sum_{i<j} Tr P_ij(x)
inline cplx plaquette_sum(const Lattice& L, const double* U) {
  double sre = 0, sim = 0;
#pragma omp parallel for reduction(+ : sre, sim) schedule(static)
  for (int64_t x = 0; x < L.ns; ++x) {
    cplx s = C(0);
    for (int i = 0; i < L.D; ++i) {
      cplx si = C(0);
      for (int j = i + 1; j < L.D; ++j) si = si + trace(pij(L, U, x, {i, true}, {j, true}));
      s = s + si;
    }
    sre += s.re;
    sim += s.im;
  }
  return {sre, sim};
}
// hamiltonian_links, state.rs:821-849: beta * sum_x sum_{i<j} (1 - Re Tr P_ij / CA)
inline double hamiltonian_links(const Lattice& L, const double* U, double beta, double CA) {
  double h = 0;
#pragma omp parallel for reduction(+ : h) schedule(static)
  for (int64_t x = 0; x < L.ns; ++x) {
    double s = 0;
    for (int i = 0; i < L.D; ++i) {
      double si = 0;
      for (int j = i + 1; j < L.D; ++j) si += 1.0 - trace(pij(L, U, x, {i, true}, {j, true})).re / CA;
      s += si;
    }
    h += s;
  }
  return h * beta;
}
// hamiltonian_efield, state.rs:1370-1385: beta * sum_x sum_i trace_squared(E_i(x))
inline double hamiltonian_efield(const Lattice& L, const double* E, double beta) {
  double h = 0;
#pragma omp parallel for reduction(+ : h) schedule(static)
  for (int64_t x = 0; x < L.ns; ++x) {
    double s = 0;
    for (int i = 0; i < L.D; ++i) s += trace_squared(E + (x * L.D + i) * 8);
    h += s;
  }
  return h * beta;
}

// ---------------------------------------------------------------- derivatives
// derivative_u, state.rs:1407-1417:  E.to_matrix() * U * (i sqrt(2 CA)) * (1/a)
inline Mat3 derivative_u(const Lattice& L, const double* U, const double* E, int64_t lidx, double CA) {
  cplx c = C(0, std::sqrt(2.0 * CA));
  Mat3 u = load3(U + lidx * 18);
  Mat3 e = adjoint_to_matrix(E + lidx * 8);
  return e * u * c * C(1.0 / L.a);
}
// integrate_link, integrator/mod.rs:216-233: first-order Euler  U + dU * dt
inline Mat3 integrate_link(const Lattice& L, const double* U, const double* E, int64_t lidx, double dt, double CA) {
  return load3(U + lidx * 18) + derivative_u(L, U, E, lidx, CA) * C(dt);
}
// Sum of adjoint staples used by the force: sum_{d in (0+,0-,1+,1-,..), |d| != i} S_{i,d}(x)^dagger
//   state.rs:1430-1438, direction order procedural_macro/src/lib.rs:56-63
inline Mat3 force_staple_sum(const Lattice& L, const double* U, int64_t x, int i) {
  Mat3 s = zero3();
  for (int j = 0; j < L.D; ++j) {
    if (j == i) continue;
    s = s + adj(sij(L, U, x, {i, true}, {j, true}));
    s = s + adj(sij(L, U, x, {i, true}, {j, false}));
  }
  return s;
}
// derivative_e for one link, state.rs:1420-1448 (LITERAL: T_a * u_i * sum_s is
// rebuilt for each of the 8 generators exactly as the reference writes it).
inline void derivative_e(const Lattice& L, const double* U, int64_t x, int i, double CA, double* out8) {
  double c = -std::sqrt(2.0 / CA);
  Mat3 u_i = link(L, U, x, {i, true});
  Mat3 sum_s = force_staple_sum(L, U, x, i);
  for (int a = 0; a < 8; ++a) out8[a] = c * trace(generator(a) * u_i * sum_s).im / L.a;
}
// Same result to rounding, 13 matmuls: hoists W = u_i * sum_s ("optimised CPU" mode of BASELINE.md §3).
inline void derivative_e_opt(const Lattice& L, const double* U, int64_t x, int i, double CA, double* out8) {
  double c = -std::sqrt(2.0 / CA);
  Mat3 w = link(L, U, x, {i, true}) * force_staple_sum(L, U, x, i);
  for (int a = 0; a < 8; ++a) out8[a] = c * trace(generator(a) * w).im / L.a;
}
// integrate_efield, integrator/mod.rs:240-254:  E + dE * dt  (whole lattice, rayon map over sites
// symplectic_euler_rayon.rs:88-104)
inline void efield_step(const Lattice& L, const double* U, const double* Ein, double* Eout, double dt, double CA,
                        bool literal = true) {
#pragma omp parallel for schedule(static)
  for (int64_t x = 0; x < L.ns; ++x)
    for (int i = 0; i < L.D; ++i) {
      double d[8];
      if (literal) derivative_e(L, U, x, i, CA, d);
      else derivative_e_opt(L, U, x, i, CA, d);
      int64_t l = x * L.D + i;
      for (int a = 0; a < 8; ++a) Eout[l * 8 + a] = Ein[l * 8 + a] + d[a] * dt;
    }
}
// link_matrix_integrate, symplectic_euler_rayon.rs:65-81
inline void link_step(const Lattice& L, const double* Uin, const double* E, double* Uout, double dt, double CA) {
#pragma omp parallel for schedule(static)
  for (int64_t l = 0; l < L.nl(); ++l) store3(Uout + l * 18, integrate_link(L, Uin, E, l, dt, CA));
}
inline void force_field(const Lattice& L, const double* U, double* F, double CA) {
#pragma omp parallel for schedule(static)
  for (int64_t x = 0; x < L.ns; ++x)
    for (int i = 0; i < L.D; ++i) derivative_e(L, U, x, i, CA, F + (x * L.D + i) * 8);
}

// The integrator compositions, symplectic_euler_rayon.rs:120-252.  U, E updated in place.
enum IntegrateKind { SYNC_SYNC = 0, LEAP_LEAP = 1, SYNC_LEAP = 2, LEAP_SYNC = 3, SYMPLECTIC = 4 };
inline void integrate(const Lattice& L, std::vector<double>& U, std::vector<double>& E, int kind, double dt, double CA,
                      bool literal = true) {
  std::vector<double> U2(U.size()), E2(E.size());
  switch (kind) {
    case SYNC_SYNC:  // both from the old state                              :127-143
      link_step(L, U.data(), E.data(), U2.data(), dt, CA);
      efield_step(L, U.data(), E.data(), E2.data(), dt, CA, literal);
      U.swap(U2);
      E.swap(E2);
      break;
    case LEAP_LEAP:  // U(dt) then E(dt) with the new U                     :145-168
      link_step(L, U.data(), E.data(), U2.data(), dt, CA);
      efield_step(L, U2.data(), E.data(), E2.data(), dt, CA, literal);
      U.swap(U2);
      E.swap(E2);
      break;
    case SYNC_LEAP:  // E(dt/2), links unchanged                           :170-191
      efield_step(L, U.data(), E.data(), E2.data(), dt / 2.0, CA, literal);
      E.swap(E2);
      break;
    case LEAP_SYNC:  // U(dt) then E(dt/2) with the new U                  :193-218
      link_step(L, U.data(), E.data(), U2.data(), dt, CA);
      efield_step(L, U2.data(), E.data(), E2.data(), dt / 2.0, CA, literal);
      U.swap(U2);
      E.swap(E2);
      break;
    case SYMPLECTIC: {  // E(dt/2) -> U(dt) with E_half -> E(dt/2) with U_new :220-252
      efield_step(L, U.data(), E.data(), E2.data(), dt / 2.0, CA, literal);
      link_step(L, U.data(), E2.data(), U2.data(), dt, CA);
      efield_step(L, U2.data(), E2.data(), E.data(), dt / 2.0, CA, literal);
      U.swap(U2);
      break;
    }
  }
}

// ---------------------------------------------------------------- Gauss law
// EField::gauss, field.rs:1174-1195:
//   G(x) = sum_i [ E_i(x) - U_{-i}(x) E_i(x-i) U_{-i}(x)^dagger ],  U_{-i}(x) = U_i(x-i)^dagger
inline Mat3 gauss(const Lattice& L, const double* U, const double* E, int64_t x) {
  Mat3 g = zero3();
  for (int i = 0; i < L.D; ++i) {
    Mat3 e_i = adjoint_to_matrix(E + (x * L.D + i) * 8);
    Mat3 u_mi = link(L, U, x, {i, false});
    int64_t p_mi = shift(L, x, {i, false});
    Mat3 e_m_i = adjoint_to_matrix(E + (p_mi * L.D + i) * 8);
    g = g + (e_i - u_mi * e_m_i * adj(u_mi));
  }
  return g;
}
// gauss_sum_div, field.rs:1199-1220: sum_x | Tr( (sum_a T_a) G(x) ) |
inline double gauss_sum_div(const Lattice& L, const double* U, const double* E) {
  Mat3 tsum = zero3();
  for (int a = 0; a < 8; ++a) tsum = tsum + generator(a);
  double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t x = 0; x < L.ns; ++x) {
    cplx t = trace(tsum * gauss(L, U, E, x));
    s += std::sqrt(norm2(t));
  }
  return s;
}
// project_to_gauss_step, field.rs:1301-1337 (K = 0.12, product U G U^dagger G_p as written)
inline void project_to_gauss_step(const Lattice& L, const double* U, const double* Ein, double* Eout) {
  const cplx K = C(0.12, 0.0);
#pragma omp parallel for schedule(static)
  for (int64_t x = 0; x < L.ns; ++x) {
    for (int i = 0; i < L.D; ++i) {
      Mat3 u = link(L, U, x, {i, true});
      Mat3 g = gauss(L, U, Ein, x);
      Mat3 gp = gauss(L, U, Ein, shift(L, x, {i, true}));
      Mat3 m = (u * g * adj(u) * gp - g) * K;
      const double* e = Ein + (x * L.D + i) * 8;
      for (int a = 0; a < 8; ++a) {
        Mat3 t = generator(a);
        Eout[(x * L.D + i) * 8 + a] = 2.0 * trace(t * (m + t * C(e[a]))).re;
      }
    }
  }
}
// project_to_gauss, field.rs:1265-1294.  Returns iterations (steps) done, or -1 on NaN.
inline int64_t project_to_gauss(const Lattice& L, const double* U, std::vector<double>& E, int64_t max_steps = 1 << 20) {
  std::vector<double> tmp(E.size());
  project_to_gauss_step(L, U, E.data(), tmp.data());
  E.swap(tmp);
  int64_t steps = 1;
  for (;;) {
    double v = gauss_sum_div(L, U, E.data());
    if (std::isnan(v)) return -1;
    if (v <= EPS * (double)(L.ns * 4 * 8 * 10)) break;
    if (steps >= max_steps) return -2;
    for (int k = 0; k < 4; ++k) {
      project_to_gauss_step(L, U, E.data(), tmp.data());
      E.swap(tmp);
      ++steps;
    }
  }
  return steps;
}

// ---------------------------------------------------------------- reprojection
// ortho_matrix_from_2_vector + create_matrix_from_2_vector, su3.rs:270-303
inline Mat3 orthonormalize(const Mat3& a) {
  cplx v1[3] = {a.m[0][0], a.m[1][0], a.m[2][0]};
  cplx v2[3] = {a.m[0][1], a.m[1][1], a.m[2][1]};
  // try_normalize(eps).unwrap_or(v1): None when norm <= eps
  double n1 = std::sqrt(norm2(v1[0]) + norm2(v1[1]) + norm2(v1[2]));
  if (n1 > EPS)
    for (auto& z : v1) z = z / n1;
  // v2 - v1 * (conj(v1) . v2)   (nalgebra dot is non-conjugating)
  cplx d = conj(v1[0]) * v2[0] + conj(v1[1]) * v2[1] + conj(v1[2]) * v2[2];
  cplx w[3];
  for (int k = 0; k < 3; ++k) w[k] = v2[k] - v1[k] * d;
  double n2 = std::sqrt(norm2(w[0]) + norm2(w[1]) + norm2(w[2]));
  if (n2 > EPS)
    for (auto& z : w) z = z / n2;
  // cross(conj v1, conj v2)
  cplx a1[3] = {conj(v1[0]), conj(v1[1]), conj(v1[2])};
  cplx b1[3] = {conj(w[0]), conj(w[1]), conj(w[2])};
  cplx cr[3] = {a1[1] * b1[2] - a1[2] * b1[1], a1[2] * b1[0] - a1[0] * b1[2], a1[0] * b1[1] - a1[1] * b1[0]};
  Mat3 r;
  for (int k = 0; k < 3; ++k) {
    r.m[k][0] = v1[k];
    r.m[k][1] = w[k];
    r.m[k][2] = cr[k];
  }
  return r;
}
// LinkMatrix::normalize, field.rs:897-901
inline void normalize_links(const Lattice& L, double* U) {
#pragma omp parallel for schedule(static)
  for (int64_t l = 0; l < L.nl(); ++l) store3(U + l * 18, orthonormalize(load3(U + l * 18)));
}

// ---------------------------------------------------------------- su3_exp_i
// su3.rs:820-855, N = 26 Cayley-Hamilton recursion; field.rs:185-205 for t() and d().
inline Mat3 su3_exp_i(const double* e) {
  static double inv_fact[26];
  static bool init = false;
  if (!init) {
    // 1 / (k! as f64); factorials are exact u128 in the reference, rounded once to f64.
    long double f = 1.0L;
    for (int k = 0; k < 26; ++k) {
      if (k > 0) f *= (long double)k;
      inv_fact[k] = 1.0 / (double)f;
    }
    init = true;
  }
  const int N_LOOP = 25;
  Mat3 m = adjoint_to_matrix(e);
  cplx q0 = C(inv_fact[N_LOOP]), q1 = C(0), q2 = C(0);
  cplx d = det(m) * C(0, 1);              // d = i det(X)
  cplx t = C(-0.5 * trace_squared(e));    // t = -1/2 Tr(X^2)
  for (int i = N_LOOP - 1; i >= 0; --i) {
    cplx q0n = C(inv_fact[i]) + d * q2;
    cplx q1n = C(0, 1) * (q0 - t * q2);
    cplx q2n = C(0, 1) * q1;
    q0 = q0n; q1 = q1n; q2 = q2n;
  }
  Mat3 diag = zero3();
  for (int k = 0; k < 3; ++k) diag.m[k][k] = q0;
  return diag + m * q1 + m * m * q2;
}

// ---------------------------------------------------------------- staple / dS
// staple, monte_carlo/mod.rs:339-362 (link (x, j); sum over positive i != j)
inline Mat3 staple(const Lattice& L, const double* U, int64_t x, int j) {
  Mat3 s = zero3();
  Dir dj{j, true};
  for (int i = 0; i < L.D; ++i) {
    if (i == j) continue;
    Dir di{i, true};
    Mat3 el_1 = adj(sij(L, U, x, dj, di));
    Mat3 u1 = link(L, U, shift(L, x, dj), -di);
    int64_t xmi = shift(L, x, -di);
    Mat3 u2 = adj(link(L, U, xmi, dj));
    Mat3 u3 = link(L, U, xmi, di);
    s = s + (el_1 + u1 * u2 * u3);
  }
  return s;
}
// delta_s_old_new_cmp, monte_carlo/mod.rs:324-334
inline double delta_s(const Mat3& stap, const Mat3& new_link, const Mat3& old_link, double beta, double CA) {
  return -trace((new_link - old_link) * stap).re * beta / CA;
}

// ---------------------------------------------------------------- RNG
// The reference draws from rand 0.8 StdRng (ChaCha12) -- un-vendored, streams unpinned.
// The oracle and the CUDA path share ONE specified generator instead:
// Philox4x32-10 (Salmon et al., SC'11), key = seed, counter = (block, link, call-counter).
struct Philox {
  static inline void round(uint32_t c[4], const uint32_t k[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  static inline void block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
      round(c, k);
      k[0] += 0x9E3779B9u;
      k[1] += 0xBB67AE85u;
    }
    for (int i = 0; i < 4; ++i) out[i] = c[i];
  }
};
// One stream per (seed, call counter, global link index); sequential 53-bit draws.
struct Stream {
  uint32_t key[2], ctr[4], buf[4];
  int have = 0;
  Stream(uint64_t seed, uint64_t counter, uint64_t idx) {
    key[0] = (uint32_t)seed;
    key[1] = (uint32_t)(seed >> 32);
    ctr[0] = 0;
    ctr[1] = (uint32_t)idx;
    ctr[2] = (uint32_t)counter;
    ctr[3] = ((uint32_t)(counter >> 32) & 0x00FFFFFFu) | (((uint32_t)(idx >> 32) & 0xFFu) << 24);
  }
  inline uint64_t bits53() {
    if (have == 0) {
      Philox::block(ctr, key, buf);
      ctr[0] += 1;
      have = 2;
    }
    int o = (2 - have) * 2;
    --have;
    return ((uint64_t)(buf[o] >> 5) << 26) | (uint64_t)(buf[o + 1] >> 6);
  }
  inline double uniform01() { return (double)bits53() * 0x1.0p-53; }             // [0,1)
  inline double open_closed01() { return (double)(bits53() + 1) * 0x1.0p-53; }   // (0,1]
  inline double uniform_pm1() { return 2.0 * uniform01() - 1.0; }               // [-1,1)
  inline bool bernoulli(double p) { return uniform01() < p; }
  // Box-Muller pair from one Philox block (two 53-bit draws).
  inline void normal_pair(double& z0, double& z1) {
    double u1 = open_closed01();
    double u2 = uniform01();
    double r = std::sqrt(-2.0 * std::log(u1));
    double th = 2.0 * PI * u2;
    z0 = r * std::cos(th);
    z1 = r * std::sin(th);
  }
};

// random_su3, su3.rs:322-355 (Gram-Schmidt of two Uniform(-1,1)^6 vectors)
template <class R>
inline Mat3 random_su3(R& rng) {
  auto rv = [&](cplx v[3]) {
    for (int k = 0; k < 3; ++k) {
      double re = rng.uniform_pm1();
      double im = rng.uniform_pm1();
      v[k] = {re, im};
    }
  };
  cplx v1[3], v2[3];
  rv(v1);
  while (std::sqrt(norm2(v1[0]) + norm2(v1[1]) + norm2(v1[2])) <= EPS) rv(v1);
  rv(v2);
  // v1.dot(&v2): non-conjugating dot                                    su3.rs:343
  while (std::sqrt(norm2(v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2])) <= EPS) rv(v2);
  Mat3 m = zero3();
  for (int k = 0; k < 3; ++k) {
    m.m[k][0] = v1[k];
    m.m[k][1] = v2[k];
  }
  return orthonormalize(m);
}
// Behaviour flags (process-global in the oracle; per-context in the CUDA library).
//   FLAG_PAULI3_FIXED: use the true sigma_3 = diag(1,-1).  Default (0) restates the reference AS CODED:
//   PAULI_3 = diag(1, 1) (su2.rs:39-45; its doc comment says diag(1,-1)), which makes every
//   complex_matrix_from_vec output non-unitary when x_3 != 0.
constexpr int FLAG_PAULI3_FIXED = 1;
//   FLAG_UNIFORM_DIRECTION: draw the direction of the heat-bath SU(2) vector uniformly on the sphere (reject cube
//   samples outside the unit ball).  Default (0) restates distribution.rs:199-219 AS CODED: a Uniform(-1,1)^3 sample
//   normalised to unit length, which over-weights the cube diagonals (not the heat-bath conditional distribution).
constexpr int FLAG_UNIFORM_DIRECTION = 16;
inline int g_flags = 0;
constexpr int KP_MAX_ITER = 10000;  // the reference loops forever on NaN parameters; both ports cap and return x0 = 1
// complex_matrix_from_vec, su2.rs:134-140
inline Mat2 complex_matrix_from_vec(double x0, const double x[3]) {
  Mat2 r;
  r.m[0][0] = C(x0, x[2]);
  r.m[0][1] = C(x[1], x[0]);
  r.m[1][0] = C(-x[1], x[0]);
  r.m[1][1] = (g_flags & FLAG_PAULI3_FIXED) ? C(x0, -x[2]) : C(x0, x[2]);
  return r;
}
// random_su2_close_to_unity, su2.rs:80-100
template <class R>
inline Mat2 random_su2_close_to_unity(double spread, R& rng) {
  double r[3];
  for (auto& v : r) v = rng.uniform_pm1();
  double n = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  double x[3];
  for (int k = 0; k < 3; ++k) x[k] = (n > EPS ? r[k] / n : r[k]) * spread;
  double x0u = std::sqrt(1.0 - (x[0] * x[0] + x[1] * x[1] + x[2] * x[2]));
  double x0 = rng.bernoulli(0.5) ? x0u : -x0u;
  return complex_matrix_from_vec(x0, x);
}
// get_r / get_s / get_t, su3.rs:428-530
inline Mat3 get_r(const Mat2& m) {
  Mat3 r = ident3();
  r.m[0][0] = m.m[0][0]; r.m[0][1] = m.m[0][1]; r.m[1][0] = m.m[1][0]; r.m[1][1] = m.m[1][1];
  return r;
}
inline Mat3 get_s(const Mat2& m) {
  Mat3 r = ident3();
  r.m[0][0] = m.m[0][0]; r.m[0][2] = m.m[0][1]; r.m[2][0] = m.m[1][0]; r.m[2][2] = m.m[1][1];
  return r;
}
inline Mat3 get_t(const Mat2& m) {
  Mat3 r = ident3();
  r.m[1][1] = m.m[0][0]; r.m[1][2] = m.m[0][1]; r.m[2][1] = m.m[1][0]; r.m[2][2] = m.m[1][1];
  return r;
}
// get_sub_block_{r,s,t}, su3.rs:559-620
inline Mat2 sub_block(const Mat3& m, int which) {
  static const int ia[3] = {0, 0, 1}, ib[3] = {1, 2, 2};
  int a = ia[which], b = ib[which];
  Mat2 r;
  r.m[0][0] = m.m[a][a]; r.m[0][1] = m.m[a][b]; r.m[1][0] = m.m[b][a]; r.m[1][1] = m.m[b][b];
  return r;
}
// project_to_su2_unorm, su2.rs:155-157:  m - m^dagger + 1 * conj(tr m)
inline Mat2 project_to_su2_unorm(const Mat2& m) {
  Mat2 a = adj(m), r;
  cplx t = conj(trace(m));
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) r.m[i][j] = m.m[i][j] - a.m[i][j] + (i == j ? t : C(0));
  return r;
}
// random_su3_close_to_unity, su3.rs:385-398
template <class R>
inline Mat3 random_su3_close_to_unity(double spread, R& rng) {
  Mat3 r = get_r(random_su2_close_to_unity(spread, rng));
  Mat3 s = get_s(random_su2_close_to_unity(spread, rng));
  Mat3 t = get_t(random_su2_close_to_unity(spread, rng));
  Mat3 x = r * s * t;
  if (rng.bernoulli(0.5)) x = adj(x);
  return x;
}
inline bool is_normal(double v) { return std::isnormal(v); }
// random_su2, su2.rs:200-216
template <class R>
inline Mat2 random_su2(R& rng) {
  cplx v[2];
  double n;
  int guard = 0;
  do {
    for (auto& z : v) {
      double re = rng.uniform_pm1();
      double im = rng.uniform_pm1();
      z = {re, im};
    }
    n = std::sqrt(norm2(v[0]) + norm2(v[1]));
  } while (!is_normal(n) && ++guard < KP_MAX_ITER);
  cplx a = v[0] / n, b = v[1] / n;
  Mat2 r;
  r.m[0][0] = a; r.m[0][1] = b; r.m[1][0] = -conj(b); r.m[1][1] = conj(a);
  return r;
}
// ModifiedNormal + HeatBathDistributionNorm (Kennedy-Pendleton), distribution.rs:89-100, 336-350
template <class R>
inline double heat_bath_norm(double param_exp, R& rng) {
  for (int it = 0; it < KP_MAX_ITER; ++it) {
    double r = rng.uniform01();
    double r0 = rng.open_closed01(), r1 = rng.open_closed01(), r2 = rng.open_closed01();
    double c = std::cos(2.0 * PI * r1);
    double lambda = std::sqrt(-(std::log(r0) + c * c * std::log(r2)) / (2.0 * param_exp));
    if (r * r <= 1.0 - lambda * lambda) return 1.0 - 2.0 * (lambda * lambda);
  }
  return 1.0;
}
// HeatBathDistribution -> SU(2)-like matrix, distribution.rs:199-219
template <class R>
inline Mat2 heat_bath_matrix(double param_exp, R& rng) {
  double x0 = heat_bath_norm(param_exp, rng);
  double xu[3], n;
  int guard = 0;
  do {
    for (auto& v : xu) v = rng.uniform_pm1();
    n = std::sqrt(xu[0] * xu[0] + xu[1] * xu[1] + xu[2] * xu[2]);
  } while ((n <= EPS || ((g_flags & FLAG_UNIFORM_DIRECTION) && n > 1.0)) && ++guard < KP_MAX_ITER);
  double sc = std::sqrt(1.0 - x0 * x0);
  double x[3] = {xu[0] / n * sc, xu[1] / n * sc, xu[2] / n * sc};
  return complex_matrix_from_vec(x0, x);
}
// heat_bath_su2, heat_bath.rs:73-86.  `coupling` = beta * coupling_scale (reference: scale 1).
template <class R>
inline Mat2 heat_bath_su2(const Mat2& stap, double coupling, R& rng) {
  double k = std::sqrt(det(stap).re);
  if (is_normal(k)) {
    Mat2 v = adj(stap);
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j) v.m[i][j] = v.m[i][j] / k;
    Mat2 x = heat_bath_matrix(coupling * k, rng);
    return x * v;
  }
  return random_su2(rng);
}
// HeatBathSweep::get_modif, heat_bath.rs:90-109
template <class R>
inline Mat3 heat_bath_link(const Mat3& u, const Mat3& a, double coupling, R& rng) {
  Mat3 r = get_r(heat_bath_su2(project_to_su2_unorm(sub_block(u * a, 0)), coupling, rng));
  Mat3 s = get_s(heat_bath_su2(project_to_su2_unorm(sub_block(r * u * a, 1)), coupling, rng));
  Mat3 t = get_t(heat_bath_su2(project_to_su2_unorm(sub_block(s * r * u * a, 2)), coupling, rng));
  return t * s * r * u;
}
// MetropolisHastingsSweep::potential_modif, metropolis_hastings_sweep.rs:126-143
template <class R>
inline Mat3 metropolis_proposal(const Mat3& old_link, int n_update, double spread, R& rng) {
  Mat3 nl = old_link;
  for (int k = 0; k < n_update; ++k) {
    Mat3 rm = orthonormalize(random_su3_close_to_unity(spread, rng));
    nl = rm * nl;
  }
  return nl;
}

// ---------------------------------------------------------------- 3x3 complex SVD
// nalgebra SVD::new(a, true, true) (overrelaxation.rs:95, 167) is un-vendored; the
// over-relaxation results are SVD-convention independent for non-degenerate singular
// values, so any accurate SVD serves: one-sided Jacobi (Hestenes).
//   a = u * diag(s) * v^dagger
inline void svd3(const Mat3& a, Mat3& u, double s[3], Mat3& v) {
  Mat3 w = a;  // columns get orthogonalised: w = a * v
  v = ident3();
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double app = 0, aqq = 0;
        cplx apq = C(0);
        for (int k = 0; k < 3; ++k) {
          app += norm2(w.m[k][p]);
          aqq += norm2(w.m[k][q]);
          apq = apq + conj(w.m[k][p]) * w.m[k][q];
        }
        double g = std::sqrt(norm2(apq));
        if (g <= 1e-300 || g <= 1e-17 * std::sqrt(app * aqq)) continue;
        off = std::fmax(off, g / std::sqrt(app * aqq));
        cplx ph = apq / g;  // e^{i phi}
        double zeta = (aqq - app) / (2.0 * g);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), sn = c * t;
        for (int k = 0; k < 3; ++k) {
          cplx wp = w.m[k][p], wq = w.m[k][q];
          w.m[k][p] = wp * c - wq * conj(ph) * sn;
          w.m[k][q] = wp * ph * sn + wq * c;
          cplx vp = v.m[k][p], vq = v.m[k][q];
          v.m[k][p] = vp * c - vq * conj(ph) * sn;
          v.m[k][q] = vp * ph * sn + vq * c;
        }
      }
    if (off < 1e-15) break;
  }
  u = zero3();
  for (int j = 0; j < 3; ++j) {
    double n = 0;
    for (int k = 0; k < 3; ++k) n += norm2(w.m[k][j]);
    n = std::sqrt(n);
    s[j] = n;
    for (int k = 0; k < 3; ++k) u.m[k][j] = (n > 0) ? w.m[k][j] / n : C(k == j ? 1.0 : 0.0);
  }
}
// su3::reverse, su3.rs:705-714
inline Mat3 reverse(const Mat3& a) {
  Mat3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = (i == j) ? a.m[i][j] : -a.m[i][j];
  return r;
}
// OverrelaxationSweepRotation::get_modif, overrelaxation.rs:86-98
inline Mat3 overrelax_rotation(const Mat3& ulink, const Mat3& stap) {
  Mat3 a = adj(stap), u, v;
  double s[3];
  svd3(a, u, s, v);
  Mat3 rot = u * adj(v);
  return rot * adj(ulink) * rot;
}
// OverrelaxationSweepReverse::get_modif, overrelaxation.rs:158-171
inline Mat3 overrelax_reverse(const Mat3& ulink, const Mat3& stap) {
  Mat3 a = adj(stap), u, v;
  double s[3];
  svd3(a, u, s, v);
  Mat3 vt = adj(v);
  return u * reverse(adj(u) * ulink * adj(vt)) * vt;
}

// ---------------------------------------------------------------- sweeps
// Visit orders.  SEQUENTIAL = the reference's Gauss-Seidel loop over get_links()
// (index order, heat_bath.rs:118-121).  CHECKERBOARD = for dir, for parity: all
// links (x, dir) with parity(x) = p -- the order the CUDA path uses; within a
// sub-step the links do not interact, so a serial loop reproduces the parallel update.
enum SweepOrder { SEQUENTIAL = 0, CHECKERBOARD = 1 };
inline int site_parity(const Lattice& L, int64_t x) {
  int64_t p = 0;
  for (int k = 0; k < L.D; ++k) p += (x / L.stride[k]) % L.ext[k];
  return (int)(p & 1);
}
// Lattices with odd extents: two colours do not decouple a periodic ring of odd length; the CUDA path uses the classes
// (boundary mask, parity), bit k of the mask set iff ext[k] is odd and x_k = ext[k] - 1 (lq_site_class).
inline int site_boundary_mask(const Lattice& L, int64_t x) {
  int m = 0;
  for (int k = 0; k < L.D; ++k)
    if ((L.ext[k] & 1) && (x / L.stride[k]) % L.ext[k] == L.ext[k] - 1) m |= 1 << k;
  return m;
}
template <class F>
inline void for_each_link(const Lattice& L, int order, F f) {
  if (order == SEQUENTIAL) {
    for (int64_t l = 0; l < L.nl(); ++l) f(l / L.D, (int)(l % L.D));
  } else {
    int om = 0;
    for (int k = 0; k < L.D; ++k)
      if (L.ext[k] & 1) om |= 1 << k;
    for (int d = 0; d < L.D; ++d)
      for (int cm = 0; cm <= om; ++cm) {
        if (cm & ~om) continue;
        for (int p = 0; p < 2; ++p)
          for (int64_t x = 0; x < L.ns; ++x)
            if (site_parity(L, x) == p && site_boundary_mask(L, x) == cm) f(x, d);
      }
  }
}
// rng mode: per_link = true  -> Stream(seed, counter, global link index) per link (CUDA-comparable)
//           per_link = false -> ONE serial stream for the whole sweep (reference style)
struct SweepRng {
  uint64_t seed, counter;
  bool per_link;
  Stream serial;
  SweepRng(uint64_t s, uint64_t c, bool pl) : seed(s), counter(c), per_link(pl), serial(s, c, 0xFFFFFFFFFFull) {}
};
// HeatBathSweep::next_element_default, heat_bath.rs:113-123
inline void sweep_heatbath(const Lattice& L, double* U, double beta, double coupling_scale, int order, SweepRng& rng) {
  for_each_link(L, order, [&](int64_t x, int d) {
    int64_t l = x * L.D + d;
    Mat3 u = load3(U + l * 18);
    Mat3 a = staple(L, U, x, d);
    if (rng.per_link) {
      Stream st(rng.seed, rng.counter, (uint64_t)l);
      store3(U + l * 18, heat_bath_link(u, a, beta * coupling_scale, st));
    } else {
      store3(U + l * 18, heat_bath_link(u, a, beta * coupling_scale, rng.serial));
    }
  });
}
// Option beyond the crate: over-relaxation inside the three SU(2) sub-groups (Brown-Woch reflections on
// the Cabibbo-Marinari blocks of heat_bath.rs:90-109).  For each block: w = 2x2 block of cur*A, m = project(w)/k in SU(2),
// left-multiply by (m^dagger)^2, which maps the block to its reflection m^dagger: Re tr unchanged, links stay in SU(3)
// (the crate's SVD variants return U(3) matrices, overrelaxation.rs:96-97).
inline Mat3 overrelax_su2(const Mat3& ulink, const Mat3& stap) {
  Mat3 cur = ulink;
  for (int which = 0; which < 3; ++which) {
    Mat2 p = project_to_su2_unorm(sub_block(cur * stap, which));
    double k = std::sqrt(det(p).re);
    if (!is_normal(k)) continue;
    Mat2 v = adj(p);
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j) v.m[i][j] = v.m[i][j] / k;
    Mat2 x = v * v;
    cur = (which == 0 ? get_r(x) : which == 1 ? get_s(x) : get_t(x)) * cur;
  }
  return cur;
}
inline Mat3 overrelax_any(const Mat3& u, const Mat3& a, int kind) {
  return kind == 0 ? overrelax_rotation(u, a) : kind == 1 ? overrelax_reverse(u, a) : overrelax_su2(u, a);
}
// OverrelaxationSweep{Rotation,Reverse}::next_element_default, overrelaxation.rs:100-110, 173-184
inline void sweep_overrelax(const Lattice& L, double* U, int kind /*0 rotation, 1 reverse, 2 SU(2) sub-groups*/, int order) {
  for_each_link(L, order, [&](int64_t x, int d) {
    int64_t l = x * L.D + d;
    Mat3 u = load3(U + l * 18);
    Mat3 a = staple(L, U, x, d);
    store3(U + l * 18, overrelax_any(u, a, kind));
  });
}
// MetropolisHastingsSweep::next_element_default, metropolis_hastings_sweep.rs:145-174
inline void sweep_metropolis(const Lattice& L, double* U, double beta, double CA, int n_update, double spread, int order,
                             SweepRng& rng, int64_t* n_accept, double* sum_prob) {
  int64_t acc = 0;
  double sp = 0;
  for_each_link(L, order, [&](int64_t x, int d) {
    int64_t l = x * L.D + d;
    Mat3 old = load3(U + l * 18);
    auto body = [&](auto& st) {
      Mat3 prop = metropolis_proposal(old, n_update, spread, st);
      Mat3 a = staple(L, U, x, d);
      double proba = std::fmax(std::fmin(std::exp(-delta_s(a, prop, old, beta, CA)), 1.0), 0.0);
      sp += proba;
      if (st.bernoulli(proba)) {
        ++acc;
        store3(U + l * 18, prop);
      }
    };
    if (rng.per_link) {
      Stream st(rng.seed, rng.counter, (uint64_t)l);
      body(st);
    } else {
      body(rng.serial);
    }
  });
  *n_accept = acc;
  *sum_prob = sp;
}
// MetropolisHastingsDeltaDiagnostic::next_element, metropolis_hastings.rs:374-417:
// `n_hits` single-link updates at uniformly random links from one serial stream.
inline void metropolis_single_link(const Lattice& L, double* U, double beta, double CA, double spread, int64_t n_hits,
                                   Stream& st, int64_t* n_accept, double* sum_prob) {
  int64_t acc = 0;
  double sp = 0;
  for (int64_t h = 0; h < n_hits; ++h) {
    int64_t x = 0;
    for (int k = 0; k < L.D; ++k) {
      int64_t c = (int64_t)(st.uniform01() * (double)L.ext[k]);
      if (c >= L.ext[k]) c = L.ext[k] - 1;
      x += c * L.stride[k];
    }
    int d = (int)(st.uniform01() * (double)L.D);
    if (d >= L.D) d = L.D - 1;
    int64_t l = x * L.D + d;
    Mat3 old = load3(U + l * 18);
    Mat3 prop = orthonormalize(random_su3_close_to_unity(spread, st)) * old;
    Mat3 a = staple(L, U, x, d);
    double proba = std::fmax(std::fmin(std::exp(-delta_s(a, prop, old, beta, CA)), 1.0), 0.0);
    sp += proba;
    if (st.bernoulli(proba)) {
      ++acc;
      store3(U + l * 18, prop);
    }
  }
  *n_accept = acc;
  *sum_prob = sp;
}

// ---------------------------------------------------------------- start configs / momenta
// LinkMatrix::new_determinist, field.rs:646-659 (random_su3 per link in index order); per-link Philox stream.
inline void links_random(const Lattice& L, double* U, uint64_t seed, uint64_t counter) {
#pragma omp parallel for schedule(static)
  for (int64_t l = 0; l < L.nl(); ++l) {
    Stream st(seed, counter, (uint64_t)l);
    store3(U + l * 18, random_su3(st));
  }
}
// EField::new_determinist with Normal(0, sigma), field.rs:1086-1099; state.rs:1097 (sigma = 0.5/beta)
inline void momenta_refresh(const Lattice& L, double* E, uint64_t seed, uint64_t counter, double sigma) {
#pragma omp parallel for schedule(static)
  for (int64_t l = 0; l < L.nl(); ++l) {
    Stream st(seed, counter, (uint64_t)l);
    for (int k = 0; k < 4; ++k) {
      double z0, z1;
      st.normal_pair(z0, z1);
      E[l * 8 + 2 * k] = sigma * z0;
      E[l * 8 + 2 * k + 1] = sigma * z1;
    }
  }
}

}  // namespace lqo
