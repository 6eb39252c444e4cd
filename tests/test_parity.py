"""Parity of the kernel path with the CPU oracle, through the C ABI (include/lqcd_b200.h).

Every test runs on two backends:
  * "cuda" (marked gpu): the product library lattice_qcd_rs_b200/liblqcd_b200.so on the B200;
  * "emu"  (CPU CI)    : the same kernel bodies compiled for the host (tests/emu.py, test infrastructure).
Tolerances: deterministic paths <= 1e-12 relative (north_star); per-link stochastic updates share the oracle's
Philox streams, so they are compared element-wise too (1e-9: a handful of libm calls sit between the streams
and the links).
"""
import numpy as np
import pytest

from oracle.oracle import Oracle
from tests.conftest import SEED_RNG

RTOL = 1e-12


@pytest.fixture(params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "emu":
        from tests import emu
        return emu.context
    import torch
    assert torch.cuda.is_available(), "the gpu-marked tests need a CUDA device (no CPU fallback)"
    from lattice_qcd_rs_b200 import Context
    return Context


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def hot(o, seed=SEED_RNG):
    return o.links_random(seed)


CASES = [(4, [4, 4, 4, 4]), (4, [8, 8, 8, 8]), (3, [6, 6, 6]), (2, [8, 8]), (4, [4, 6, 2, 8]), (4, [5, 5, 5, 5]),
         (3, [3, 4, 5])]


@pytest.mark.parametrize("D,ext", CASES)
def test_roundtrip_and_observables(backend, D, ext):
    o = Oracle(D, ext, a=1.3, beta=6.0)
    c = backend(D, ext, a=1.3, beta=6.0)
    U = hot(o)
    E = o.momenta_refresh(SEED_RNG, 1)
    c.links_upload(U)
    c.efield_upload(E)
    assert np.array_equal(c.links_download(), U)
    assert np.array_equal(c.efield_download(), E)
    ps, po = c.plaquette_sum(), o.plaquette_sum(U)
    assert abs(ps - po) <= RTOL * abs(po)
    assert abs(c.average_trace_plaquette() - o.average_trace_plaquette(U)) <= RTOL * abs(po)
    assert abs(c.hamiltonian_links() - o.hamiltonian_links(U)) <= RTOL * abs(o.hamiltonian_links(U))
    assert abs(c.hamiltonian_efield() - o.hamiltonian_efield(E)) <= RTOL * abs(o.hamiltonian_efield(E))
    assert abs(c.hamiltonian_total() - o.hamiltonian_total(U, E)) <= RTOL * abs(o.hamiltonian_total(U, E))


def test_errors(backend):
    from lattice_qcd_rs_b200 import LqError
    c = backend(4, 4)
    with pytest.raises(LqError) as e:  # StateInitializationError::IncompatibleSize, state.rs:784-786
        c.links_upload(np.zeros((c.nl - 1, 18)))
    assert e.value.name == "LQ_E_SIZE"
    with pytest.raises(LqError) as e:
        c.efield_upload(np.zeros((c.nl + 4, 8)))
    assert e.value.name == "LQ_E_SIZE"
    with pytest.raises(LqError) as e:  # LatticeCyclic::new: dim >= 2 (lattice.rs:190-201)
        backend(4, [4, 4, 1, 4])
    assert e.value.name == "LQ_E_BADARG"
    with pytest.raises(LqError):
        backend(5, [4] * 5)
    with pytest.raises(LqError) as e:  # MultiIntegrationError::ZeroIntegration, state.rs:480-482
        c.symplectic_n(0.1, 0)
    assert e.value.name == "LQ_E_ZERO_STEPS"
    with pytest.raises(LqError) as e:
        c.restore()
    assert e.value.name == "LQ_E_NOSNAPSHOT"
    odd = backend(4, [5, 4, 4, 4])  # odd extents are fine on a single rank (colour classes, test_sweeps_on_odd_extents)
    odd.links_set_cold()
    odd.sweep_heatbath(1, 0)
    with pytest.raises(LqError):  # spread must be in (0,1): metropolis_hastings_sweep.rs:73-80
        c.sweep_metropolis(1, 0, spread=1.5)


def test_cold_start_exact(backend):
    """test_sim_cold (test/mod.rs:431-453): U = 1, E = 0 is an exact fixed point of every composition."""
    c = backend(4, 4, a=1.0, beta=1.0)
    o = Oracle(4, 4)
    c.links_set_cold()
    c.efield_set_zero()
    assert np.array_equal(c.links_download(), o.cold_links())
    for kind in (2, 1, 3, 0, 4):
        c.integrate(kind, 0.1)
    c.symplectic_n(0.1, 3)
    c.leapfrog_n(0.1, 3)
    assert np.array_equal(c.links_download(), o.cold_links())
    assert np.array_equal(c.efield_download(), o.cold_efield())
    assert c.plaquette_sum() == complex(3.0 * 6 * 256, 0.0)
    assert c.hamiltonian_links() == 0.0 and c.hamiltonian_efield() == 0.0
    assert c.t == 4 + 3 + 3  # sync_leap does not advance t (symplectic_euler_rayon.rs:170-191)


@pytest.mark.parametrize("D,ext", CASES)
def test_staples_and_force(backend, D, ext):
    o = Oracle(D, ext, a=0.7, beta=6.0)
    c = backend(D, ext, a=0.7, beta=6.0)
    U = hot(o)
    c.links_upload(U)
    assert rel(c.staples(), o.staples(U)) <= RTOL
    assert rel(c.force(), o.force(U, literal=True)) <= RTOL


@pytest.mark.parametrize("D,ext", [(4, [4, 4, 4, 4]), (3, [6, 6, 6]), (4, [4, 6, 2, 8]), (4, [5, 5, 5, 5])])
def test_md_steps(backend, D, ext):
    o = Oracle(D, ext, a=1.0, beta=6.0)
    c = backend(D, ext, a=1.0, beta=6.0)
    U = hot(o)
    E = o.momenta_refresh(SEED_RNG, 2, sigma=0.7)
    c.links_upload(U)
    c.efield_upload(E)
    c.efield_step(0.01)
    E1 = o.efield_step(U, E, 0.01)
    assert rel(c.efield_download(), E1) <= RTOL
    c.link_step(0.01)
    U1 = o.link_step(U, E1, 0.01)
    assert rel(c.links_download(), U1) <= RTOL
    c.links_upload(U)
    c.link_step(0.05, use_exp=True)
    assert rel(c.links_download(), o.link_step_exp(U, E1, 0.05)) <= RTOL
    for kind, name in enumerate(["sync_sync", "leap_leap", "sync_leap", "leap_sync", "symplectic"]):
        c.links_upload(U)
        c.efield_upload(E)
        c.integrate(kind, 0.01)
        Uo, Eo = o.integrate(U, E, name, 0.01)
        assert rel(c.links_download(), Uo) <= RTOL, name
        assert rel(c.efield_download(), Eo) <= RTOL, name


def test_symplectic_trajectory_and_merge_equivalence(backend):
    """n-step trajectory (state.rs:470-492): fused + merged-kick schedule is BIT-identical to the literal
    3-kernel-per-step schedule, and both match the oracle to 1e-12."""
    from lattice_qcd_rs_b200 import FLAG_NO_KICK_MERGE
    D, ext = 4, [4, 4, 4, 4]
    o = Oracle(D, ext, a=1.0, beta=6.0)
    U = hot(o)
    E = o.momenta_refresh(SEED_RNG, 3)
    out = []
    for flags in (0, FLAG_NO_KICK_MERGE):
        c = backend(D, ext, a=1.0, beta=6.0)
        c.set_flags(flags)
        c.links_upload(U)
        c.efield_upload(E)
        c.symplectic_n(0.01, 10)
        assert c.t == 10
        out.append((c.links_download(), c.efield_download(), c.kernel_launches))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert out[0][2] < out[1][2]
    Uo, Eo = o.integrate(U, E, "symplectic", 0.01, n=10)
    assert rel(out[0][0], Uo) <= RTOL and rel(out[0][1], Eo) <= RTOL
    c = backend(D, ext, a=1.0, beta=6.0)
    c.links_upload(U)
    c.efield_upload(E)
    c.leapfrog_n(0.01, 5)
    Uo, Eo = o.leapfrog_n(U, E, 0.01, 5)
    assert rel(c.links_download(), Uo) <= RTOL and rel(c.efield_download(), Eo) <= RTOL


def test_leap_frog_energy_conservation(backend):
    """test_leap_frog (test/mod.rs:679-698): 4^4, a=1000, beta=1: |dH| < 1e-5 after one leapfrog step dt=0.01."""
    o = Oracle(4, 4, a=1000.0, beta=1.0)
    c = backend(4, 4, a=1000.0, beta=1.0)
    U = hot(o, 0)
    c.links_upload(U)
    c.momenta_refresh(0, 0)
    c.gauss_project()
    h0 = c.hamiltonian_total()
    c.leapfrog_n(0.01, 1)
    assert abs(c.hamiltonian_total() - h0) < 1e-5


def test_reunitarize(backend):
    o = Oracle(4, 4)
    c = backend(4, 4)
    rng = np.random.default_rng(5)
    U = rng.uniform(-2, 2, (o.nl, 18))
    U[0] = 0.0  # the `norm <= eps: leave unchanged` branch of su3.rs:284-294
    U[1] = 0.0
    U[1][0] = 2.0
    c.links_upload(U)
    c.reunitarize()
    got, want = c.links_download(), o.normalize_links(U)
    assert rel(got, want) <= RTOL
    assert np.array_equal(got[0], np.zeros(18))
    assert np.array_equal(got[1], np.eye(18)[0])
    # drift off SU(3) along a trajectory, then reproject: matches the oracle doing the same
    U = hot(o)
    E = o.momenta_refresh(1, 1, sigma=1.0)
    c.links_upload(U)
    c.efield_upload(E)
    c.symplectic_n(0.05, 4)
    c.reunitarize()
    Uo, _ = o.integrate(U, E, "symplectic", 0.05, n=4)
    assert rel(c.links_download(), o.normalize_links(Uo)) <= 1e-11


@pytest.mark.parametrize("D,ext", [(4, [4, 4, 4, 4]), (3, [6, 4, 8])])
def test_start_configs_and_momenta(backend, D, ext):
    o = Oracle(D, ext, beta=6.0)
    c = backend(D, ext, beta=6.0)
    c.links_set_random(SEED_RNG, 7)
    assert rel(c.links_download(), o.links_random(SEED_RNG, 7)) <= 1e-13
    c.momenta_refresh(SEED_RNG, 9)
    Eo = o.momenta_refresh(SEED_RNG, 9)
    assert np.abs(c.efield_download() - Eo).max() <= 1e-14
    assert abs(Eo.std() - 0.5 / 6.0) < 0.01  # sigma = 0.5/beta (state.rs:1097)


@pytest.mark.parametrize("D,ext", [(4, [4, 4, 4, 4]), (3, [6, 6, 6]), (4, [4, 6, 2, 8])])
def test_gauss(backend, D, ext):
    o = Oracle(D, ext, a=1.0, beta=2.0)
    c = backend(D, ext, a=1.0, beta=2.0)
    U = hot(o)
    E = o.momenta_refresh(SEED_RNG, 4)
    c.links_upload(U)
    c.efield_upload(E)
    assert rel(c.gauss_field(), o.gauss_field(U, E)) <= RTOL
    assert abs(c.gauss_sum_div() - o.gauss_sum_div(U, E)) <= RTOL * o.gauss_sum_div(U, E)
    E1 = o.project_to_gauss_step(U, E)
    for flags in (0, 4):  # two passes (default) and LQ_FLAG_GAUSS_FUSED (one fused kernel per iteration): same results
        c.set_flags(flags)
        c.efield_upload(E)
        c.gauss_project_step()
        assert rel(c.efield_download(), E1) <= RTOL
        # the Gauss field the fused kernel left behind is the Gauss field of the projected E
        assert rel(c.gauss_field(), o.gauss_field(U, E1)) <= RTOL
        c.gauss_project_step()
        assert rel(c.efield_download(), o.project_to_gauss_step(U, E1)) <= RTOL
    c.set_flags(0)
    c.efield_upload(E)
    it = c.gauss_project()
    Eo, ito = o.project_to_gauss(U, E)
    assert it == ito
    assert rel(c.efield_download(), Eo) <= 1e-10
    assert c.gauss_sum_div() <= np.finfo(float).eps * o.ns * 320 * 1.0000001


@pytest.mark.parametrize("D,ext", [(4, [4, 4, 4, 4]), (3, [6, 6, 6]), (2, [8, 8]), (4, [4, 6, 2, 8])])
def test_sweeps_match_oracle_checkerboard(backend, D, ext):
    o = Oracle(D, ext, a=1.0, beta=6.0)
    c = backend(D, ext, a=1.0, beta=6.0)
    U = hot(o)
    # heat bath (heat_bath.rs:73-123), reference coupling beta*k
    c.links_upload(U)
    c.sweep_heatbath(SEED_RNG, 11)
    Uo = o.sweep_heatbath(U, SEED_RNG, 11, order=1, per_link=True)
    assert rel(c.links_download(), Uo) <= 1e-9
    # Wilson-correct coupling scale
    c.links_upload(U)
    c.sweep_heatbath(SEED_RNG, 12, coupling_scale=1.0 / 3.0)
    assert rel(c.links_download(), o.sweep_heatbath(U, SEED_RNG, 12, coupling_scale=1.0 / 3.0)) <= 1e-9
    # over-relaxation (overrelaxation.rs:86-98, 158-171)
    for kind in (0, 1):
        c.links_upload(U)
        c.sweep_overrelax(kind)
        assert rel(c.links_download(), o.sweep_overrelax(U, kind, order=1)) <= 1e-9
    # Metropolis (metropolis_hastings_sweep.rs:126-174) with its diagnostics
    for n_update, spread in ((1, 0.1), (3, 0.25)):
        c.links_upload(U)
        na, sp = c.sweep_metropolis(SEED_RNG, 13, spread=spread, n_update=n_update)
        Uo, nao, spo = o.sweep_metropolis(U, SEED_RNG, 13, n_update=n_update, spread=spread, order=1, per_link=True)
        assert na == nao and abs(sp - spo) <= 1e-9 * max(spo, 1.0)
        assert rel(c.links_download(), Uo) <= 1e-9


def test_pauli3_flag(backend):
    from lattice_qcd_rs_b200 import FLAG_PAULI3_FIXED
    o = Oracle(4, 4, beta=6.0)
    c = backend(4, 4, beta=6.0)
    U = hot(o)
    try:
        o.set_flags(Oracle.FLAG_PAULI3_FIXED)
        c.set_flags(FLAG_PAULI3_FIXED)
        c.links_upload(U)
        c.sweep_heatbath(SEED_RNG, 21)
        assert rel(c.links_download(), o.sweep_heatbath(U, SEED_RNG, 21)) <= 1e-9
    finally:
        o.set_flags(0)


def test_overrelax_conserves_action(backend):
    """same_energy_reverse / same_energy_rotation (overrelaxation.rs:220-253)."""
    o = Oracle(3, 4, beta=1.0)
    c = backend(3, 4, beta=1.0)
    c.links_upload(hot(o))
    h = c.hamiltonian_links()
    for kind in (0, 1):
        c.sweep_overrelax(kind)
        h2 = c.hamiltonian_links()
        assert abs(h - h2) < np.finfo(float).eps * 100 * 4 ** 3 * (h + h2) * 0.5 * 10
        h = h2


def test_delta_s_equals_delta_h(backend):
    """test_mh_delta (metropolis_hastings.rs:480-514): staple-based dS == H_links(new) - H_links(old)."""
    o = Oracle(4, 4, beta=2.0)
    c = backend(4, 4, beta=2.0)
    U = hot(o)
    c.links_upload(U)
    A = c.staples()
    h0 = c.hamiltonian_links()
    from oracle.oracle import to_c
    for l in (0, 17, 333, o.nl - 1):
        new = o.links_random(99)[l]
        ds = o.delta_s(to_c(A[l])[0], to_c(new)[0], to_c(U[l])[0])
        V = U.copy()
        V[l] = new
        c.links_upload(V)
        assert abs(ds - (c.hamiltonian_links() - h0)) < 1e-8


def test_hmc_trajectory(backend):
    """HybridMonteCarloDiagnostic::next_element (hybrid_monte_carlo.rs:465-471, 573-613) vs the oracle, from the
    same start configuration and the same Philox momenta."""
    o = Oracle(4, 4, a=1.0, beta=6.0)
    c = backend(4, 4, a=1.0, beta=6.0)
    U = hot(o)
    c.links_upload(U)
    for k in range(3):
        r = c.hmc_trajectory(0.01, 10, SEED_RNG, k)
        ro = o.hmc_trajectory(U, 0.01, 10, SEED_RNG, k)
        assert r["gauss_steps"] == ro["gauss_steps"]
        assert abs(r["h_old"] - ro["h_old"]) <= 1e-11 * abs(ro["h_old"])
        assert abs(r["h_new"] - ro["h_new"]) <= 1e-11 * abs(ro["h_new"])
        assert abs(r["prob"] - ro["prob"]) <= 1e-6
        assert r["accepted"] == ro["accepted"]
        U = ro["U"]
        assert rel(c.links_download(), U) <= 1e-10
    # reject path: absurd step size -> prob 0 -> links restored bit-exactly
    before = c.links_download()
    r = c.hmc_trajectory(1.0, 3, SEED_RNG, 99)
    assert not r["accepted"] and r["prob"] < 1e-6 and np.isfinite(r["h_new"])
    assert np.array_equal(c.links_download(), before)


def test_snapshot_restore(backend):
    o = Oracle(3, 4, beta=6.0)
    c = backend(3, 4, beta=6.0)
    U, E = hot(o), o.momenta_refresh(1, 2)
    c.links_upload(U)
    c.efield_upload(E)
    c.set_t(5)
    c.snapshot()
    c.symplectic_n(0.02, 3)
    assert c.t == 8 and not np.array_equal(c.links_download(), U)
    c.restore()
    assert c.t == 5 and np.array_equal(c.links_download(), U) and np.array_equal(c.efield_download(), E)


@pytest.mark.parametrize("ext", [[4, 4, 4, 4], [32, 4, 2, 2]])
def test_pipelined_marshalling(backend, ext):
    """lq_links_upload_begin / _commit, lq_links_download_begin, lq_copies_wait: the same bytes as the synchronous calls
    (LatticeStateNew::new, link_matrix(), state.rs:779-815), with the next upload begun before the previous download ends."""
    o = Oracle(4, ext, a=1.0, beta=6.0)
    c = backend(4, ext, a=1.0, beta=6.0)
    inputs = [o.links_random(SEED_RNG, k) for k in range(3)]
    outs = [np.empty_like(inputs[0]) for _ in range(3)]
    E = o.momenta_refresh(SEED_RNG, 2)
    ref = []
    for U in inputs:  # the serial sequence
        c.links_upload(U)
        c.efield_upload(E)
        c.symplectic_n(0.01, 2)
        ref.append(c.links_download().copy())
    c.links_upload_begin(inputs[0])
    for k in range(3):
        c.links_upload_commit()
        if k + 1 < 3:
            c.links_upload_begin(inputs[k + 1])  # travels while step k computes
        c.efield_upload(E)
        c.symplectic_n(0.01, 2)
        c.links_download_begin(outs[k])
    c.copies_wait()
    for k in range(3):
        assert np.array_equal(outs[k], ref[k])
    with pytest.raises(Exception):
        c.links_upload_commit()  # nothing begun


@pytest.mark.parametrize("ext", [[8, 8, 8, 8], [4, 6, 2, 8], [32, 4, 4, 2], [32, 8, 2, 2], [6, 4, 4, 4]])
def test_gauss_iteration_variants_agree(backend, ext):
    """project_to_gauss (field.rs:1265-1337) through both iteration forms -- projection-step kernel then Gauss-field
    kernel (default), and the one-pass functor that recomputes the backward neighbours (LQ_FLAG_GAUSS_FUSED).  Same
    arithmetic in the same order: E, the Gauss field left behind and the iteration count must agree with each other
    (to the contraction choices of two different kernels) and with the oracle."""
    from lattice_qcd_rs_b200 import FLAG_GAUSS_FUSED
    from lattice_qcd_rs_b200._capi import FLAG_GAUSS_TWO_PASS
    o = Oracle(4, ext, a=1.0, beta=6.0)
    U = hot(o)
    E = o.momenta_refresh(SEED_RNG, 9)
    Eo, ito = o.project_to_gauss(U, E)
    out = []
    # 0: the default -- on the CUDA library lq_gauss_project iterates on the transported field U^+ E U (one kernel per
    # iteration, lq_gausst4_kernel); TWO_PASS / FUSED: the two older iteration forms
    for flags in (0, FLAG_GAUSS_TWO_PASS, FLAG_GAUSS_FUSED):
        c = backend(4, ext, a=1.0, beta=6.0)
        c.set_flags(flags)
        c.links_upload(U)
        c.efield_upload(E)
        c.gauss_project_step()
        e1, g1 = c.efield_download(), c.gauss_field()
        c.gauss_project_step()
        c.gauss_project_step()
        e3, g3 = c.efield_download(), c.gauss_field()
        c.efield_upload(E)
        it = c.gauss_project()
        assert it == ito
        ef = c.efield_download()
        assert rel(ef, Eo) <= 1e-10
        out.append((e1, g1, e3, g3, ef))
    assert rel(out[0][0], o.project_to_gauss_step(U, E)) <= RTOL
    for other in out[1:]:
        for a, b in zip(out[0], other):
            assert rel(a, b) <= 1e-13


@pytest.mark.parametrize("ext", [[8, 8, 8, 8], [4, 6, 2, 8], [32, 4, 4, 2]])
def test_tuned_kernels_equal_generic_kernels(backend, ext):
    """D = 4 has hand-tuned kernels for the MD loop and the heat-bath / over-relaxation sub-steps
    (csrc/lq_tuned.cuh); LQ_FLAG_GENERIC_KERNELS selects the dimension-generic functors instead.  Both must agree
    with each other (to rounding of the staple summation order) and with the oracle."""
    from lattice_qcd_rs_b200 import FLAG_GENERIC_KERNELS
    o = Oracle(4, ext, a=1.0, beta=6.0)
    U = hot(o)
    E = o.momenta_refresh(SEED_RNG, 7)
    Uo, Eo = o.integrate(U, E, "symplectic", 0.01, n=3)
    Uhb = o.sweep_heatbath(U, SEED_RNG, 31)
    Uor = o.sweep_overrelax(U, 1)
    Umh, nam, spm = o.sweep_metropolis(U, SEED_RNG, 41, n_update=2, spread=0.1, order=1, per_link=True)
    Egs = o.project_to_gauss_step(U, E)
    res = []
    for flags in (0, FLAG_GENERIC_KERNELS):
        c = backend(4, ext, a=1.0, beta=6.0)
        c.set_flags(flags)
        c.links_upload(U)
        c.efield_upload(E)
        # plaquette reduction (tuned: lq_plaq4_kernel): same terms, another summation order
        pq, hl = c.average_trace_plaquette(), c.hamiltonian_links()
        assert abs(pq - o.average_trace_plaquette(U)) <= RTOL * abs(pq)
        assert abs(hl - o.hamiltonian_links(U)) <= RTOL * abs(hl)
        c.symplectic_n(0.01, 3)
        md = (c.links_download(), c.efield_download())
        assert rel(md[0], Uo) <= RTOL and rel(md[1], Eo) <= RTOL
        c.links_upload(U)
        c.sweep_heatbath(SEED_RNG, 31)
        hb = c.links_download()
        assert rel(hb, Uhb) <= 1e-9
        c.links_upload(U)
        c.sweep_overrelax(1)
        orx = c.links_download()
        assert rel(orx, Uor) <= 1e-9
        # Gauss projection step (tuned: lq_gstep4_kernel)
        c.links_upload(U)
        c.efield_upload(E)
        c.gauss_project_step()
        gs = c.efield_download()
        assert rel(gs, Egs) <= RTOL
        c.links_upload(U)
        na, sp = c.sweep_metropolis(SEED_RNG, 41, spread=0.1, n_update=2)
        assert na == nam and abs(sp - spm) <= 1e-9 * spm and rel(c.links_download(), Umh) <= 1e-9
        res.append((md, hb, orx, pq, hl, gs))
    assert rel(res[0][0][0], res[1][0][0]) <= 1e-14 and rel(res[0][0][1], res[1][0][1]) <= 1e-14
    assert rel(res[0][1], res[1][1]) <= 1e-12 and rel(res[0][2], res[1][2]) <= 1e-12
    assert abs(res[0][3] - res[1][3]) <= 1e-14 * abs(res[1][3]) and abs(res[0][4] - res[1][4]) <= 1e-13 * abs(res[1][4])
    assert rel(res[0][5], res[1][5]) <= 1e-14


@pytest.mark.parametrize("D,ext", [(4, [4, 4, 4, 4]), (3, [4, 6, 4])])
def test_integrator_options(backend, D, ext):
    """Integrators beyond the crate's (lq_set_integrator / lq_md_n): compositions of the reference's own
    E and U updates.  Parity against the oracle's composition of the same steps; with the exponential link update the
    links stay in SU(3) and the map is time-reversible; Omelyan's energy error beats leap-frog's at equal step."""
    from lattice_qcd_rs_b200 import INTEGRATOR_OMELYAN, INTEGRATOR_SYMPLECTIC_EULER, OMELYAN_LAMBDA
    o = Oracle(D, ext, a=1.0, beta=6.0)
    c = backend(D, ext, a=1.0, beta=6.0)
    U = hot(o)
    E = o.momenta_refresh(SEED_RNG, 21)
    h0 = o.hamiltonian_total(U, E)
    dh = {}
    for kind, use_exp in ((INTEGRATOR_SYMPLECTIC_EULER, False), (INTEGRATOR_SYMPLECTIC_EULER, True),
                          (INTEGRATOR_OMELYAN, False), (INTEGRATOR_OMELYAN, True)):
        c.links_upload(U)
        c.efield_upload(E)
        c.set_t(0)
        c.set_integrator(kind, OMELYAN_LAMBDA, use_exp)
        c.md_n(0.02, 5)
        Uo, Eo = o.md_n(U, E, 0.02, 5, kind=kind, lam=OMELYAN_LAMBDA, use_exp=use_exp)
        Ug, Eg = c.links_download(), c.efield_download()
        assert c.t == 5
        assert rel(Ug, Uo) <= RTOL and rel(Eg, Eo) <= RTOL
        dh[(kind, use_exp)] = abs(c.hamiltonian_total() - h0)
        if use_exp:
            # unitarity is kept to rounding (the Euler update drifts at O(dt^2) per step) ...
            M = Ug.reshape(-1, 3, 3, 2)
            M = (M[..., 0] + 1j * M[..., 1]).transpose(0, 2, 1)  # column-major AoS -> row-major matrices
            assert np.abs(M @ M.conj().transpose(0, 2, 1) - np.eye(3)).max() <= 1e-13
            # ... and the step is reversible: flip the momenta, integrate back, recover the start
            c.efield_upload(-Eg)
            c.md_n(0.02, 5)
            assert rel(c.links_download(), U) <= 1e-11 and rel(-c.efield_download(), E) <= 1e-11
    assert dh[(INTEGRATOR_OMELYAN, True)] < 0.5 * dh[(INTEGRATOR_SYMPLECTIC_EULER, True)]
    # the default selection is the reference's integrator: lq_md_n == lq_symplectic_n
    c.set_integrator(INTEGRATOR_SYMPLECTIC_EULER, OMELYAN_LAMBDA, False)
    c.links_upload(U)
    c.efield_upload(E)
    c.md_n(0.01, 3)
    a = (c.links_download(), c.efield_download())
    c.links_upload(U)
    c.efield_upload(E)
    c.symplectic_n(0.01, 3)
    assert np.array_equal(a[0], c.links_download()) and np.array_equal(a[1], c.efield_download())
    from lattice_qcd_rs_b200 import LqError
    with pytest.raises(LqError):
        c.set_integrator(INTEGRATOR_OMELYAN, 0.7, True)
    with pytest.raises(LqError):
        c.set_integrator(5, 0.2, True)


@pytest.mark.parametrize("D,ext", [(4, [5, 4, 3, 4]), (3, [3, 5, 7]), (2, [5, 6]), (4, [3, 3, 3, 3])])
def test_sweeps_on_odd_extents(backend, D, ext):
    """The reference's sequential sweeps accept any N >= 2 (lattice.rs:190-201).  On a periodic ring of odd length two
    colours do not decouple the links of a sub-step, so lattices with odd extents are swept in the colour classes
    (boundary mask, parity) of lq_site_class -- restated in the oracle's checkerboard order; element-wise parity as for
    even lattices, and every link is visited exactly once per sweep."""
    o = Oracle(D, ext, a=1.0, beta=6.0)
    c = backend(D, ext, a=1.0, beta=6.0)
    U = hot(o)
    c.links_upload(U)
    c.sweep_heatbath(SEED_RNG, 11)
    got = c.links_download()
    assert rel(got, o.sweep_heatbath(U, SEED_RNG, 11, order=1, per_link=True)) <= 1e-9
    assert np.all(np.abs(got - U).reshape(o.nl, 18).max(axis=1) > 1e-6)  # no link left out
    for kind in (0, 1, 2):
        c.links_upload(U)
        h0 = c.hamiltonian_links()
        c.sweep_overrelax(kind)
        assert rel(c.links_download(), o.sweep_overrelax(U, kind, order=1)) <= 1e-9
        assert abs(c.hamiltonian_links() - h0) <= 1e-10 * abs(h0)  # microcanonical only if no two links interfered
    c.links_upload(U)
    na, sp = c.sweep_metropolis(SEED_RNG, 13, spread=0.1, n_update=2)
    Uo, nao, spo = o.sweep_metropolis(U, SEED_RNG, 13, n_update=2, spread=0.1, order=1, per_link=True)
    assert na == nao and abs(sp - spo) <= 1e-9 * spo and rel(c.links_download(), Uo) <= 1e-9
