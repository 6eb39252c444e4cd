"""Pins the CPU oracle against every known-answer test / invariant the reference holds for the hot path
(SURVEY.md section 4 / 8c).  Each test names the reference test it restates.  CPU only."""
import numpy as np
import pytest
import scipy.linalg

from oracle.oracle import Oracle, from_c, to_c
from tests.conftest import SEED_RNG
from tests.np_ref import NpLattice, dag, gell_mann_half

EPS = 1e-9  # the reference tests' EPSILON (field.rs:1454, su3.rs:1014, test/mod.rs:17)


def cm(rows):
    return np.array(rows, dtype=np.complex128)


# ---------------------------------------------------------------- Philox known answers (Random123 kat_vectors)
def test_philox4x32_10_known_answers():
    o = Oracle(2, 2)
    assert o.philox_block([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert o.philox_block([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert o.philox_block([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_stream_uniform_moments_and_independence():
    o = Oracle(2, 2)
    u = o.stream_uniform01(SEED_RNG, 3, 17, 200000)
    assert 0.0 <= u.min() and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 4 / np.sqrt(12 * u.size)
    assert abs(u.var() - 1 / 12) < 1e-3
    v = o.stream_uniform01(SEED_RNG, 3, 18, 200000)
    assert abs(np.corrcoef(u, v)[0, 1]) < 0.01
    w = o.stream_uniform01(SEED_RNG, 4, 17, 1000)
    assert not np.allclose(u[:1000], w)


# ---------------------------------------------------------------- test_generators (test/mod.rs:234-250), test_gen (su3.rs:1160-1177)
def test_generators():
    o = Oracle(2, 2)
    T = [o.generator(a) for a in range(8)]
    ref = gell_mann_half()
    for a in range(8):
        assert np.array_equal(T[a], ref[a]) or np.allclose(T[a], ref[a], atol=1e-16)
        assert np.trace(T[a]) == 0 or abs(np.trace(T[a])) < 1e-16
        assert np.array_equal(T[a], dag(T[a]))
        if a < 7:
            assert abs(np.linalg.det(T[a])) == 0
        for b in range(8):
            assert abs(abs(np.trace(T[a] @ T[b])) - 0.5 * (a == b)) < 1e-15
    # Su3Adjoint::to_matrix doctest (field.rs:98-105): e = (1,0,..) -> GENERATORS[0]
    e = np.zeros(8)
    e[0] = 1
    assert np.array_equal(o.adjoint_to_matrix(e), T[0])


# ---------------------------------------------------------------- test_exp_basic (test/mod.rs:95-146), equivalence_exp_i (:148-167)
@pytest.mark.parametrize("factor", [1.0, 2.0, -1.254, 4.254])
def test_exp_basic(factor):
    o = Oracle(2, 2)
    c, s = np.cos(factor), np.sin(factor)
    e = np.zeros(8)
    e[0] = 2 * factor
    assert np.allclose(o.su3_exp_i(e), cm([[c, 1j * s, 0], [1j * s, c, 0], [0, 0, 1]]), atol=EPS, rtol=0)
    e = np.zeros(8)
    e[1] = 2 * factor
    assert np.allclose(o.su3_exp_i(e), cm([[c, s, 0], [-s, c, 0], [0, 0, 1]]), atol=EPS, rtol=0)


def test_equivalence_exp_i_and_su3_property():
    o = Oracle(2, 2)
    rng = np.random.default_rng(SEED_RNG % 2**32)
    T = gell_mann_half()
    vs = [np.eye(8)[i] for i in range(8)] + [rng.uniform(-np.pi, np.pi, 8) for _ in range(100)]
    for v in vs:
        m = o.su3_exp_i(v)
        ref = scipy.linalg.expm(1j * np.einsum("a,aij->ij", v, T))
        assert np.allclose(m, ref, atol=EPS, rtol=0)
        # su3_property (test/mod.rs:252-261): det = 1, unitary
        assert abs(np.linalg.det(m) - 1) < EPS and np.allclose(m @ dag(m), np.eye(3), atol=EPS)


# ---------------------------------------------------------------- othonomralization (test/mod.rs:455-592)
def test_orthonormalization_known_answers():
    o = Oracle(2, 2)
    Z, I = np.zeros((3, 3), complex), np.eye(3, dtype=complex)
    assert np.array_equal(o.orthonormalize(Z), Z)
    assert np.array_equal(o.orthonormalize(I), I)
    m = cm([[1, 0, 0], [0, 0, 0], [0, 0, 0]])
    assert np.array_equal(o.orthonormalize(m), m)
    assert np.array_equal(o.orthonormalize(cm([[2, 0, 0], [0, 2, 0], [0, 0, 0]])), I)
    assert np.array_equal(o.orthonormalize(cm([[2j, 0, 0], [0, 2j, 0], [0, 0, 0]])), np.diag([1j, 1j, -1]))
    assert np.array_equal(o.orthonormalize(cm([[0, 1, 0], [1, 0, 0], [0, 0, 0]])),
                          cm([[0, 1, 0], [1, 0, 0], [0, 0, -1]]))
    assert np.array_equal(o.orthonormalize(cm([[0, 1, 0], [0, 0, 0], [1, 0, 0]])),
                          cm([[0, 1, 0], [0, 0, 1], [1, 0, 0]]))
    rng = np.random.default_rng(1)
    for _ in range(200):
        m = rng.uniform(-10, 10, (3, 3)) + 1j * rng.uniform(-10, 10, (3, 3))
        r = o.orthonormalize(m)
        assert abs(np.linalg.det(r) - 1) < EPS and np.allclose(r @ dag(r), np.eye(3), atol=EPS)


# ---------------------------------------------------------------- lattice index contract (lattice.rs:62-83, 255-323, 909-929)
def test_index_contract_and_periodic_shift():
    o = Oracle(4, 4)
    # link_canonical doctest: point [1,0,2,0], XNeg -> [0,0,2,0] XPos ; YNeg -> [1,3,2,0] YPos
    U = np.arange(o.nl * 18, dtype=np.float64).reshape(o.nl, 18)
    x = o.site_index([1, 0, 2, 0])
    assert x == 1 + 2 * 16
    # sij with (i=+y, j=-x) starts with U_{-x}(x) = U_x(x - x_hat)^dagger: check via pij/sij on a tagged config
    cold = o.cold_links()
    tag = cold.copy()
    lidx = o.site_index([0, 0, 2, 0]) * 4 + 0
    tag[lidx] = from_c(2j * np.eye(3))[0]
    s = o.sij(tag, x, Oracle.sdir(1), Oracle.sdir(0, False))  # U_{-x}(x) U_y(x-x) U_{-x}^+(x+y)
    assert np.allclose(s, dag(2j * np.eye(3)))
    lidx = o.site_index([1, 3, 2, 0]) * 4 + 1
    tag = cold.copy()
    tag[lidx] = from_c(3j * np.eye(3))[0]
    s = o.sij(tag, x, Oracle.sdir(0), Oracle.sdir(1, False))
    assert np.allclose(s, dag(3j * np.eye(3)))
    # add_point_direction_n doctest: [1,2,2,0] - 3 y_hat -> [1,1,2,0] with dim 4 (three single shifts here)
    o2 = Oracle(4, 4)
    tag = cold.copy()
    tag[o2.site_index([1, 1, 2, 0]) * 4 + 2] = from_c(5 * np.eye(3))[0]
    # U_z at x-3y == U_z at x+y (period 4)
    assert o2.site_index([1, 2 - 3, 2, 0]) == o2.site_index([1, 3, 2, 0])


# ---------------------------------------------------------------- magnetic_field (field.rs:1580-1711)
def test_magnetic_field_known_answers():
    o = Oracle(3, 4, a=1.0)
    I3 = np.eye(3)
    U = o.cold_links()
    X, Y, Zd = Oracle.sdir(0), Oracle.sdir(1), Oracle.sdir(2)
    assert np.allclose(o.clover(U, 0, X, Y), 4 * I3, atol=EPS)
    assert np.allclose(o.f_mu_nu(U, 0, X, Y), 0, atol=EPS)
    for d in range(3):
        assert np.allclose(o.magnetic_field(U, 0, d), 0, atol=EPS)
    U[0] = from_c(1j * I3)[0]
    assert np.allclose(o.clover(U, 0, X, Y), 2 * I3, atol=EPS)
    assert np.allclose(o.clover(U, 0, Y, X), 2 * I3, atol=EPS)
    assert np.allclose(o.f_mu_nu(U, 0, X, Y), 0, atol=EPS)
    for d in range(3):
        assert np.allclose(o.magnetic_field(U, 0, d), 0, atol=EPS)
    U = o.cold_links()
    U[o.site_index([1, 0, 0]) * 3 + 1] = from_c(1j * I3)[0]
    assert np.allclose(o.clover(U, 0, X, Y), (3 + 1j) * I3, atol=EPS)
    assert np.allclose(o.clover(U, 0, Y, X), (3 - 1j) * I3, atol=EPS)
    assert np.allclose(o.f_mu_nu(U, 0, X, Y), 0.25j * I3, atol=EPS)
    assert np.allclose(o.magnetic_field(U, 0, 0), 0, atol=EPS)
    assert np.allclose(o.magnetic_field(U, 0, 1), 0, atol=EPS)
    assert np.allclose(o.magnetic_field(U, 0, 2), 0.25 * I3, atol=EPS)
    # point [4,0,0] == [0,0,0] on a period-4 lattice
    assert o.site_index([4, 0, 0]) == 0


# ---------------------------------------------------------------- independent numpy stencils vs oracle
@pytest.mark.parametrize("D,n", [(4, 4), (3, 6), (2, 8)])
def test_oracle_vs_numpy_stencils(D, n):
    o = Oracle(D, n, a=0.7, beta=2.0)
    U = o.links_random(SEED_RNG, 1)
    # drift the links off SU(3) so that no unitarity shortcut can hide
    U += 1e-3 * np.random.default_rng(0).normal(size=U.shape)
    E = np.random.default_rng(1).normal(size=(o.nl, 8))
    npl = NpLattice(D, [n] * D, a=0.7)
    ps = o.plaquette_sum(U)
    assert abs(ps - npl.plaquette_sum(U)) < 1e-12 * abs(ps)
    assert np.allclose(to_c(o.staples(U)), npl.staple_mc(U), rtol=0, atol=1e-12)
    F = o.force(U, literal=True)
    assert np.allclose(F, npl.force(U), rtol=0, atol=1e-12)
    assert np.allclose(F, o.force(U, literal=False), rtol=0, atol=1e-13)
    assert np.allclose(to_c(o.gauss_field(U, E)), npl.gauss(U, E), rtol=0, atol=1e-12)
    # hamiltonians from their definitions
    npl_h = o.beta * ((o.ns * D * (D - 1) // 2) - ps.real / 3.0)
    assert abs(o.hamiltonian_links(U) - npl_h) < 1e-11 * abs(npl_h)
    assert abs(o.hamiltonian_efield(E) - o.beta * 0.5 * (E**2).sum()) < 1e-11 * (E**2).sum()


def test_anisotropic_extents():
    ext = [4, 2, 6, 8]
    o = Oracle(4, ext, a=1.3, beta=1.0)
    U = o.links_random(7, 0)
    npl = NpLattice(4, ext, a=1.3)
    assert abs(o.plaquette_sum(U) - npl.plaquette_sum(U)) < 1e-10
    assert np.allclose(o.force(U), npl.force(U), rtol=0, atol=1e-12)


# ---------------------------------------------------------------- derivative_u / integrate_link (state.rs:1407-1417, integrator/mod.rs:216-233)
def test_link_step_formula():
    o = Oracle(3, 4, a=2.0)
    U = o.links_random(3, 0)
    E = np.random.default_rng(2).normal(size=(o.nl, 8))
    dt = 0.013
    T = gell_mann_half()
    Em = np.einsum("na,aij->nij", E, T)
    ref = to_c(U) + dt * (1j * np.sqrt(6.0) / 2.0) * (Em @ to_c(U))
    assert np.allclose(to_c(o.link_step(U, E, dt)), ref, rtol=0, atol=1e-14)
    # exponential variant agrees with scipy expm
    Ux = to_c(o.link_step_exp(U, E, dt))
    for n in range(0, o.nl, 37):
        assert np.allclose(Ux[n], scipy.linalg.expm(1j * dt * np.sqrt(6.0) / 2.0 * Em[n]) @ to_c(U)[n], atol=1e-12)


# ---------------------------------------------------------------- test_sim_cold (test/mod.rs:431-453): exact fixed point
def test_sim_cold_exact_fixed_point():
    o = Oracle(4, 10, a=10.0, beta=0.1)
    U, E = o.cold_links(), o.cold_efield()
    U2, E2 = o.integrate(U, E, "sync_leap", 0.1)
    assert np.array_equal(U, U2) and np.array_equal(E, E2)
    U3, E3 = o.integrate(U2, E2, "leap_leap", 0.1)
    assert np.array_equal(U2, U3) and np.array_equal(E2, E3)
    U4, E4 = o.integrate(U3, E3, "leap_sync", 0.1)
    assert np.array_equal(U3, U4) and np.array_equal(E3, E4)


# ---------------------------------------------------------------- test_sim_hamiltonian / test_gauss_law (test/mod.rs:327-427)
def test_sim_hamiltonian_and_gauss_law():
    rng = np.random.default_rng(5)
    o = Oracle(4, 6, a=100.0, beta=1.0)  # reference uses 10^4; 6^4 keeps the CPU suite fast
    U = o.links_random(SEED_RNG, 0)
    E = rng.uniform(-np.pi, np.pi, size=(o.nl, 8))
    h = o.hamiltonian_total(U, E)
    U2, E2 = o.integrate(U, E, "sync_sync", 1e-4)
    assert h - o.hamiltonian_total(U2, E2) < 0.01
    o = Oracle(4, 6, a=1.0, beta=1.0)
    U2, E2 = o.integrate(U, E, "sync_sync", 1e-6)
    g1, g2 = to_c(o.gauss_field(U, E)), to_c(o.gauss_field(U2, E2))
    assert np.abs(g1 - g2).max() < 1e-3


# ---------------------------------------------------------------- test_leap_frog (test/mod.rs:679-698)
def test_leap_frog_energy():
    o = Oracle(4, 4, a=1000.0, beta=1.0)
    U = o.links_random(0, 0)
    E, it = o.project_to_gauss(U, o.momenta_refresh(0, 1))
    assert it > 0
    h1 = o.hamiltonian_total(U, E)
    U2, E2 = o.leapfrog_n(U, E, 0.01, 1)
    assert abs(h1 - o.hamiltonian_total(U2, E2)) < 1e-5


# ---------------------------------------------------------------- integrator (test/integrator.rs:10-67)
def test_integrator_sequences():
    DT = 1e-4
    o = Oracle(3, 4, a=1.0, beta=8.0)
    U = o.cold_links()
    for k in range(10):
        U, _, _ = o.sweep_metropolis(U, SEED_RNG, k, n_update=1, spread=0.1, order=0, per_link=False)
    E, it = o.project_to_gauss(U, o.momenta_refresh(SEED_RNG, 100))
    assert it > 0
    h = o.hamiltonian_total(U, E)
    U2, E2 = o.integrate(U, E, "symplectic", DT, n=10)
    assert abs(h - o.hamiltonian_total(U2, E2)) < 1e-4
    Ul, El = o.integrate(U, E, "sync_leap", DT)
    Ul, El = o.integrate(Ul, El, "leap_leap", DT, n=1)
    Ul, El = o.integrate(Ul, El, "leap_sync", DT)
    Ul, El = o.integrate(Ul, El, "sync_sync", DT, n=1)
    assert abs(h - o.hamiltonian_total(Ul, El)) < 1e-5
    U3, E3 = o.leapfrog_n(U, E, DT, 10)
    assert abs(h - o.hamiltonian_total(U3, E3)) < 1e-5
    # merged half steps == literal symplectic steps to rounding (what the GPU's fused trajectory relies on)
    assert np.allclose(U3, U2, rtol=0, atol=1e-14) and np.allclose(E3, E2, rtol=0, atol=1e-14)


# ---------------------------------------------------------------- test_mh_delta (metropolis_hastings.rs:480-514)
def test_delta_s_equals_delta_h():
    o = Oracle(3, 4, a=1.0, beta=1.0)
    U = o.links_random(SEED_RNG, 0)
    st = to_c(o.staples(U))
    rng = np.random.default_rng(3)
    for _ in range(10):
        l = int(rng.integers(o.nl))
        old = to_c(U)[l]
        new = o.random_su3_close_to_unity(SEED_RNG, 9, l, 0.1) @ old
        U2 = U.copy()
        U2[l] = from_c(new)[0]
        ds = o.delta_s(st[l], new, old)
        assert abs(np.exp(-ds) - np.exp(o.hamiltonian_links(U) - o.hamiltonian_links(U2))) < 1e-8


# ---------------------------------------------------------------- SVD + over-relaxation (overrelaxation.rs:86-98, 158-171, 220-253)
def test_svd3_and_overrelax_vs_numpy():
    o = Oracle(2, 2)
    rng = np.random.default_rng(11)
    for _ in range(50):
        a = rng.normal(size=(3, 3)) + 1j * rng.normal(size=(3, 3))
        u, s, v = o.svd3(a)
        assert np.allclose(u @ np.diag(s) @ dag(v), a, atol=1e-13)
        assert np.allclose(u @ dag(u), np.eye(3), atol=1e-13) and np.allclose(v @ dag(v), np.eye(3), atol=1e-13)
        assert np.allclose(np.sort(s), np.sort(np.linalg.svd(a, compute_uv=False)), atol=1e-13)
        # same over-relaxed link from LAPACK's SVD (convention independence)
        ulink = o.orthonormalize(rng.normal(size=(3, 3)) + 1j * rng.normal(size=(3, 3)))
        un, sn, vhn = np.linalg.svd(dag(a))
        rot = un @ vhn
        assert np.allclose(o.overrelax_link(ulink, a, 0), rot @ dag(ulink) @ rot, atol=1e-12)
        w = dag(un) @ ulink @ dag(vhn)
        rev = np.where(np.eye(3, dtype=bool), w, -w)
        assert np.allclose(o.overrelax_link(ulink, a, 1), un @ rev @ vhn, atol=1e-12)


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("order", [0, 1])
def test_overrelax_same_energy(kind, order):
    """same_energy_rotation / same_energy_reverse: |H - H'| < eps * 100 * 4^3 * mean(H, H')."""
    o = Oracle(3, 4, a=1.0, beta=1.0)
    U = o.links_random(SEED_RNG, 0)
    h = o.hamiltonian_links(U)
    U2 = o.sweep_overrelax(U, kind, order=order)
    h2 = o.hamiltonian_links(U2)
    assert abs(h - h2) < np.finfo(float).eps * 100 * 4**3 * (h + h2) * 0.5


# ---------------------------------------------------------------- su2 (su2.rs:257-308), distributions (distribution.rs:459-530)
def test_su2_projection_and_kp_sampler():
    o = Oracle(2, 2)
    rng = np.random.default_rng(4)
    for _ in range(100):
        r = rng.uniform(-1, 1, (2, 2)) + 1j * rng.uniform(-1, 1, (2, 2))
        inp = np.stack([r.real, r.imag], -1).reshape(8).copy()
        out = np.empty(8)
        import ctypes as C
        o.L.lqo_project_to_su2_unorm(inp.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)))
        p = (out.reshape(2, 2, 2)[..., 0] + 1j * out.reshape(2, 2, 2)[..., 1])
        assert abs(np.trace(p).imag) < EPS
        assert np.allclose(p @ dag(p), np.eye(2) * np.linalg.det(p), atol=EPS)
        assert np.allclose(p, r - dag(r) + np.eye(2) * np.conj(np.trace(r)))
    # Kennedy-Pendleton: density of x0 is sqrt(1-x0^2) exp(alpha x0) on [-1, 1]
    alpha = 3.0
    x = o.heat_bath_norm_samples(SEED_RNG, 0, alpha, 200000)
    assert x.min() >= -1 and x.max() <= 1
    grid = np.linspace(-1, 1, 20001)
    w = np.sqrt(1 - grid**2) * np.exp(alpha * grid)
    mean_exact = (grid * w).sum() / w.sum()
    assert abs(x.mean() - mean_exact) < 5 * x.std() / np.sqrt(x.size)


# ---------------------------------------------------------------- proposal generator quirk (su2.rs:39-45, 134-140)
def test_close_to_unity_uses_pauli3_as_coded():
    """PAULI_3 = diag(1,1) in the reference => r*s*t is NOT exactly unitary; error ~ spread."""
    o = Oracle(2, 2)
    devs = []
    for i in range(50):
        m = o.random_su3_close_to_unity(SEED_RNG, 0, i, 0.1)
        devs.append(np.abs(m @ dag(m) - np.eye(3)).max())
        fixed = o.orthonormalize(m)
        assert np.allclose(fixed @ dag(fixed), np.eye(3), atol=1e-12)
    assert 1e-4 < max(devs) < 0.5


# ---------------------------------------------------------------- sweeps: checkerboard == sequential statistics; determinism
def test_checkerboard_vs_sequential_heatbath_statistics():
    """Reference order (sequential, one serial stream) and CUDA order (checkerboard, per-link streams)
    sample the same distribution: plaquette within 2 sigma (sigma = sqrt(var/len), statistics/mod.rs:401-405).
    Needs the true sigma_3 (FLAG_PAULI3_FIXED): with PAULI_3 as coded the heat bath has no stationary
    distribution at all (next test)."""
    o = Oracle(4, 4, a=1.0, beta=2.0)  # heat bath coupling is beta*k (reference quirk) => beta_eff = 6
    o.set_flags(Oracle.FLAG_PAULI3_FIXED)
    try:
        res = {}
        for order, per_link in ((0, False), (1, True)):
            U = o.cold_links()
            vals = []
            for k in range(60):
                U = o.sweep_heatbath(U, SEED_RNG + order, k, order=order, per_link=per_link)
                if k >= 20:
                    vals.append(o.average_trace_plaquette(U).real / 3.0)
            M = to_c(U)
            assert np.allclose(M @ dag(M), np.eye(3), atol=1e-10)  # a true heat bath stays in SU(3)
            v = np.array(vals)
            res[order] = (v.mean(), v.std(ddof=1) / np.sqrt(v.size))
    finally:
        o.set_flags(0)
    (m0, s0), (m1, s1) = res[0], res[1]
    assert abs(m0 - m1) < 2.0 * 2.0 * np.hypot(s0, s1)  # factor 2: integrated autocorrelation allowance
    assert 0.3 < m0 < 0.8


def test_reference_heatbath_as_coded_leaves_su3():
    """Finding: with PAULI_3 = diag(1,1) (su2.rs:39-45) HeatBathDistribution returns non-unitary 2x2 blocks,
    so the reference's HeatBathSweep drives the links off SU(3) (|det| decays, unitarity error O(1) after a few
    sweeps).  The oracle restates that literally (flags = 0); statistical parity is only meaningful with the
    fixed sigma_3."""
    o = Oracle(4, 4, a=1.0, beta=2.0)
    U = o.cold_links()
    for k in range(5):
        U = o.sweep_heatbath(U, 1, k, order=0, per_link=False)
    M = to_c(U)
    assert np.abs(M @ dag(M) - np.eye(3)).max() > 0.1


def test_metropolis_sweep_diagnostics_and_determinism():
    o = Oracle(3, 4, a=1.0, beta=6.0)
    U = o.cold_links()
    U1, na, sp = o.sweep_metropolis(U, 5, 0, n_update=1, spread=0.1, order=1, per_link=True)
    U2, na2, sp2 = o.sweep_metropolis(U, 5, 0, n_update=1, spread=0.1, order=1, per_link=True)
    assert np.array_equal(U1, U2) and na == na2 and sp == sp2
    assert 0 < na <= o.nl and 0 < sp / o.nl <= 1.0
    U3, _, _ = o.sweep_metropolis(U, 5, 1, n_update=1, spread=0.1, order=1, per_link=True)
    assert not np.array_equal(U1, U3)
    # links stay in SU(3) to rounding because proposals are re-orthonormalised (metropolis_hastings_sweep.rs:135)
    M = to_c(U1)
    assert np.allclose(M @ dag(M), np.eye(3), atol=1e-12)


def test_hmc_trajectory_accepts_small_dt():
    o = Oracle(4, 4, a=1.0, beta=6.0)
    U = o.links_random(SEED_RNG, 0)
    r = o.hmc_trajectory(U, dt=1e-4, n_steps=5, seed=1, counter=0)
    assert r["gauss_steps"] > 0
    assert abs(r["h_old"] - r["h_new"]) < 1e-3 and r["prob"] > 0.99
    # momenta are N(0, 0.5/beta) before projection (state.rs:1097)
    E = o.momenta_refresh(1, 0)
    assert abs(E.std() - 0.5 / 6.0) < 0.01 * 0.5 / 6.0 * 5
