"""The reference's own tests for this path, replayed through the host-side mirror of its trait surface
(lattice_qcd_rs_b200/state.py: LatticeStateDefault, LatticeStateEFSyncDefault, SymplecticEulerCuda,
HybridMonteCarloDiagnostic, HeatBathSweep, ...).  Each test names the reference test / doc example it follows.

Backends: "cuda" (gpu-marked: the product library) and "emu" (CPU CI: same kernel bodies compiled for the host,
tests/emu.py -- test infrastructure).
"""
import numpy as np
import pytest

from lattice_qcd_rs_b200 import state as lq
from oracle.oracle import Oracle
from tests.conftest import SEED_RNG


@pytest.fixture(params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def lib(request):
    if request.param == "emu":
        from tests import emu
        return emu.lib()
    import torch
    assert torch.cuda.is_available(), "the gpu-marked tests need a CUDA device (no CPU fallback)"
    return None  # state.py loads the CUDA library itself


def test_lattice_and_state_errors(lib):
    """LatticeCyclic::new (lattice.rs:190-201), LatticeStateNew::new size check (state.rs:784-786)."""
    with pytest.raises(lq.LatticeInitializationError):
        lq.LatticeCyclic(0.0, 4)
    with pytest.raises(lq.LatticeInitializationError):
        lq.LatticeCyclic(float("nan"), 4)
    with pytest.raises(lq.LatticeInitializationError):
        lq.LatticeCyclic(1.0, 1)
    lat = lq.LatticeCyclic.new(1.0, 4)
    assert lat.number_of_points() == 256 and lat.number_of_canonical_links_space() == 1024  # test_iterator_length
    with pytest.raises(lq.StateInitializationError) as e:
        lq.LatticeStateDefault.new(lat, 1.0, np.zeros((1023, 18)), lib=lib)
    assert e.value.kind == "IncompatibleSize"
    with pytest.raises(lq.StateInitializationError):
        lq.LatticeStateDefault.new_cold(1.0, 1.0, 1, lib=lib)
    st = lq.LatticeStateDefault.new_cold(1.0, 1.0, 4, lib=lib)
    with pytest.raises(AssertionError):  # set_link_matrix panics on a wrong length, state.rs:808-815
        st.set_link_matrix(np.zeros((10, 18)))
    assert lq.MetropolisHastingsSweep.new(0, 0.1, lq.Rng(1)) is None       # metropolis_hastings_sweep.rs:73-80
    assert lq.MetropolisHastingsSweep.new(1, 1.0, lq.Rng(1)) is None
    with pytest.raises(lq.MultiIntegrationError):                            # state.rs:480-482
        lq.LatticeStateEFSyncDefault.new_cold(1.0, 1.0, 4, lib=lib).simulate_symplectic_n(lq.SymplecticEulerCuda(), 0.1, 0)


def test_sim_cold(lib):
    """test_sim_cold, test/mod.rs:431-453: the cold state is an exact fixed point of sync->leap->leap->sync."""
    size, number_of_pts, beta = 10.0, 4, 0.1
    sim1 = lq.LatticeStateEFSyncDefault.new_cold(size, beta, number_of_pts, lib=lib)
    integ = lq.SymplecticEulerCuda.new()
    sim2 = sim1.simulate_to_leapfrog(integ, 0.1)
    assert np.array_equal(sim1.e_field(), sim2.e_field()) and np.array_equal(sim1.link_matrix(), sim2.link_matrix())
    sim3 = sim2.simulate_leap(integ, 0.1)
    assert np.array_equal(sim2.e_field(), sim3.e_field()) and np.array_equal(sim2.link_matrix(), sim3.link_matrix())
    sim4 = sim3.simulate_to_synchronous(integ, 0.1)
    assert np.array_equal(sim3.e_field(), sim4.e_field()) and np.array_equal(sim3.link_matrix(), sim4.link_matrix())
    assert sim4.t() == 2 and sim2.t() == 0  # t + 1 except sync_leap (symplectic_euler_rayon.rs:170-191)
    assert abs(sim4.average_trace_plaquette() - 3.0) < 1e-15


def test_sim_hamiltonian_and_gauss_law(lib):
    """test_sim_hamiltonian_rayon / test_gauss_law_rayon, test/mod.rs:379-427."""
    rng = lq.Rng.seed_from_u64(SEED_RNG)
    state = lq.LatticeStateEFSyncDefault.new_determinist(100.0, 1.0, 6, rng, lib=lib)
    h = state.hamiltonian_total()
    state2 = state.simulate_sync(lq.SymplecticEulerCuda(), 1e-4)
    assert abs(h - state2.hamiltonian_total()) < 0.01
    g1 = state.gauss()
    state3 = state.simulate_sync(lq.SymplecticEulerCuda(), 1e-6)
    g2 = state3.gauss()
    assert np.sqrt(((g1 - g2) ** 2).reshape(-1, 18).sum(axis=1)).max() < 1e-3
    # the integrator returned NEW states: the original is untouched (state.rs:383-392 takes &self)
    assert state.t() == 0 and state2.t() == 1
    assert abs(state.hamiltonian_total() - h) == 0.0


def test_leap_frog(lib):
    """test_leap_frog, test/mod.rs:679-698: |dH| < 1e-5 after one leapfrog step dt = 0.01 (4^4, a = 1000, beta = 1)."""
    rng = lq.Rng.seed_from_u64(0)
    state = lq.LatticeStateEFSyncDefault.new_determinist(1000.0, 1.0, 4, rng, lib=lib)
    h = state.hamiltonian_total()
    leap = state.simulate_to_leapfrog(lq.SymplecticEulerCuda(), 0.01)
    state2 = leap.simulate_to_synchronous(lq.SymplecticEulerCuda(), 0.01)
    assert abs(h - state2.hamiltonian_total()) < 1e-5
    state3 = state.simulate_using_leapfrog_n(lq.SymplecticEulerCuda(), 0.01, 1)
    assert np.array_equal(state3.link_matrix(), state2.link_matrix())
    state4 = state.simulate_using_leapfrog_n_auto(lq.SymplecticEulerCuda(), 0.01, 3)
    assert state4.t() == 3 and abs(h - state4.hamiltonian_total()) < 1e-4


def test_integrator(lib):
    """`integrator`, test/integrator.rs:10-67: D = 3, 4^3, beta = 8; Metropolis sweeps, then symplectic steps and
    mixed leap/sync sequences conserve H."""
    rng = lq.Rng.seed_from_u64(SEED_RNG)
    state = lq.LatticeStateDefault.new_determinist(1000.0, 8.0, 4, rng, D=3, lib=lib)
    mh = lq.MetropolisHastingsSweep.new(1, 0.1, rng)
    for _ in range(10):
        state = state.monte_carlo_step(mh)
    assert 0.0 < mh.prob_replace_mean() <= 1.0 and 0 <= mh.number_replace_last() <= 3 * 64
    state.normalize_link_matrices()
    integ = lq.SymplecticEulerCuda()
    st = lq.LatticeStateEFSyncDefault.new_random_e_state(state, rng)
    h = st.hamiltonian_total()
    st2 = st.simulate_symplectic_n_auto(integ, 1e-4, 10)
    assert abs(h - st2.hamiltonian_total()) < 1e-4
    st3 = st.simulate_to_leapfrog(integ, 1e-4).simulate_leap_n(integ, 1e-4, 2).simulate_to_synchronous(integ, 1e-4)
    assert abs(h - st3.hamiltonian_total()) < 1e-5
    st4 = st.simulate_symplectic(integ, 1e-4)
    st5 = st.simulate_symplectic_n(integ, 1e-4, 1)
    assert np.array_equal(st4.link_matrix(), st5.link_matrix()) and np.array_equal(st4.e_field(), st5.e_field())


def test_hmc_like_the_doc_example(lib):
    """hybrid_monte_carlo.rs doc example (:23-52) + config 1 of BASELINE.json (8^4 shrunk to 4^4 on the CPU CI):
    HybridMonteCarloDiagnostic + symplectic Euler, trajectories compared with the oracle from the same start
    configuration and the same Philox momenta."""
    n = 4
    rng = lq.Rng.seed_from_u64(SEED_RNG)
    state = lq.LatticeStateDefault.new_determinist(1.0, 6.0, n, rng, lib=lib)
    o = Oracle(4, n, a=1.0, beta=6.0)
    U = np.array(state.link_matrix())
    hmc = lq.HybridMonteCarloDiagnostic.new(0.01, 10, lq.SymplecticEulerCuda.new(), rng)
    probe = lq.Rng(rng.state)  # replays the (seed, counter) pairs the method will draw
    for _ in range(3):
        seed, counter = probe.next_u64(), probe.next_u64() >> 8
        state = state.monte_carlo_step(hmc)
        ro = o.hmc_trajectory(U, 0.01, 10, seed, counter)
        assert hmc.has_replace_last() == ro["accepted"]
        assert abs(hmc.prob_replace_last() - ro["prob"]) <= 1e-6
        assert hmc.gauss_steps_last == ro["gauss_steps"]
        U = ro["U"]
        assert np.abs(state.link_matrix() - U).max() <= 1e-10
    p = state.average_trace_plaquette().real / 3.0
    assert abs(p - o.average_trace_plaquette(U).real / 3.0) <= 1e-12
    assert isinstance(state, lq.LatticeStateDefault)  # next_element returns the plain link state (state_owned)


def test_sweeps_through_monte_carlo_trait(lib):
    """heat_bath.rs / overrelaxation.rs doc examples: state.monte_carlo_step(&mut method) in a loop;
    over-relaxation conserves the action (same_energy_rotation/reverse, overrelaxation.rs:220-253)."""
    rng = lq.Rng.seed_from_u64(SEED_RNG)
    state = lq.LatticeStateDefault.new_determinist(1.0, 2.0, 4, rng, lib=lib)
    hb = lq.HeatBathSweep.new(rng)
    p0 = state.average_trace_plaquette().real / 3.0
    for _ in range(5):
        state = state.monte_carlo_step(hb)
    p1 = state.average_trace_plaquette().real / 3.0
    assert p1 > p0 + 0.1  # a hot start orders under the heat bath
    for method in (lq.OverrelaxationSweepReverse.new(), lq.OverrelaxationSweepRotation.new()):
        h = state.hamiltonian_links()
        state = state.monte_carlo_step(method)
        assert abs(h - state.hamiltonian_links()) <= 1e-10 * abs(h)
    combo = lq.HybridMethodVec([hb, lq.OverrelaxationSweepReverse.new()])
    state = state.monte_carlo_step(combo)
    state.normalize_link_matrices()
    assert np.isfinite(state.hamiltonian_links())


def test_omelyan_integrator_through_the_trait_surface(lib):
    """OmelyanCuda (an option the crate does not have) plugs into the same places as the reference's
    integrators: simulate_symplectic(_n) and HybridMonteCarloDiagnostic::new(delta_t, n, integrator, rng)."""
    rng = lq.Rng(SEED_RNG)
    st = lq.LatticeStateEFSyncDefault.new_determinist(1.0, 6.0, 4, rng, lib=lib)
    h0 = st.hamiltonian_total()
    euler = st.simulate_symplectic_n(lq.SymplecticEulerCuda.new(), 0.02, 10)
    omel = st.simulate_symplectic_n(lq.OmelyanCuda.new(), 0.02, 10)
    one_by_one = st
    for _ in range(2):
        one_by_one = one_by_one.simulate_symplectic(lq.OmelyanCuda.new(), 0.02)
    two = st.simulate_symplectic_n(lq.OmelyanCuda.new(), 0.02, 2)
    assert one_by_one.t() == two.t() == 2
    # n merged steps == n single steps up to the rounding of the merged kick (2 l dt vs l dt + l dt)
    assert np.abs(one_by_one.link_matrix() - two.link_matrix()).max() <= 1e-13
    assert omel.t() == euler.t() == 10
    assert abs(omel.hamiltonian_total() - h0) < 0.2 * abs(euler.hamiltonian_total() - h0)
    # st itself is untouched (integrators return new states) and still integrates with the reference's rule
    again = st.simulate_symplectic_n(lq.SymplecticEulerCuda.new(), 0.02, 10)
    assert np.array_equal(again.link_matrix(), euler.link_matrix())
    plain = lq.LatticeStateDefault.new(st.lattice(), 6.0, st.link_matrix(), lib=lib)
    hmc = lq.HybridMonteCarloDiagnostic(0.02, 10, lq.OmelyanCuda.new(), lq.Rng(5))
    plain = plain.monte_carlo_step(hmc)
    p_om = hmc.prob_replace_last()
    plain2 = lq.LatticeStateDefault.new(st.lattice(), 6.0, st.link_matrix(), lib=lib)
    hmc2 = lq.HybridMonteCarloDiagnostic(0.02, 10, lq.SymplecticEulerCuda.new(), lq.Rng(5))
    plain2 = plain2.monte_carlo_step(hmc2)
    assert 0.0 <= hmc2.prob_replace_last() <= 1.0 and 0.0 <= p_om <= 1.0
    assert p_om >= hmc2.prob_replace_last() - 1e-12  # same momenta (same seed), smaller energy error


def test_monte_carlo_default_and_wrapper(lib):
    """MonteCarloDefault + McWrapper (monte_carlo/mod.rs:117-293, doc example :81-116): MetropolisHastingsDiagnostic
    wrapped with a generator, stepped through LatticeState::monte_carlo_step on a cold 4^3 state at beta = 6."""
    rng = lq.Rng(0)
    assert lq.MetropolisHastingsDiagnostic.new(0, 0.1) is None and lq.MetropolisHastingsDiagnostic.new(1, 1.5) is None
    mh = lq.MetropolisHastingsDiagnostic.new(1, 0.1)
    wrapper = lq.McWrapper.new(mh, rng)
    state = lq.LatticeStateDefault.new_cold(1.0, 6.0, 4, D=3, lib=lib)
    h0 = state.hamiltonian_links()
    assert h0 == 0.0
    n_acc = 0
    for _ in range(30):
        old = state
        state = state.monte_carlo_step(wrapper)
        p = mh.prob_replace_last()
        assert 0.0 <= p <= 1.0
        if mh.has_replace_last():
            n_acc += 1
            assert state is not old
            # probability_of_replacement is exp(H_old - H_new) clamped (monte_carlo/mod.rs:137-142)
            want = min(1.0, np.exp(old.hamiltonian_links() - state.hamiltonian_links()))
            assert abs(p - want) <= 1e-12
        else:
            assert state is old
    assert 0 < n_acc <= 30 and state.hamiltonian_links() > 0.0
    mcd, rng2 = wrapper.deconstruct()
    assert mcd is mh and rng2 is rng and wrapper.mcd() is mh and wrapper.rng_mut() is rng


def test_hybrid_method_couple(lib):
    """HybridMethodCouple / Triple (hybrid.rs:330-446): methods applied in order, errors tagged with their position."""
    rng = lq.Rng(SEED_RNG)
    state = lq.LatticeStateDefault.new_determinist(1.0, 6.0, 4, rng, lib=lib)
    hb = lq.HeatBathSweep.new(lq.Rng(1))
    ov = lq.OverrelaxationSweepReverse.new()
    couple = lq.HybridMethodCouple.new(hb, ov)
    assert couple.method_1() is hb and couple.method_2() is ov and couple.deconstruct() == (hb, ov)
    # same generator state => the couple equals the two calls made by hand
    ref = state.clone()
    ref = lq.HeatBathSweep.new(lq.Rng(1)).next_element(ref)
    ref = ov.next_element(ref)
    state = state.monte_carlo_step(couple)
    assert np.array_equal(state.link_matrix(), ref.link_matrix())
    triple = lq.HybridMethodTriple(hb, ov, lq.OverrelaxationSweepRotation.new())
    h = state.hamiltonian_links()
    state = triple.next_element(state)
    assert np.isfinite(state.hamiltonian_links()) and state.hamiltonian_links() != h

    class Boom(lq.MonteCarlo):
        def next_element(self, state):
            raise ValueError("boom")

    with pytest.raises(lq.HybridMethodCoupleError) as e:
        lq.HybridMethodCouple(Boom(), ov).next_element(state)
    assert e.value.which == "ErrorFirst" and isinstance(e.value.error, ValueError)
    with pytest.raises(lq.HybridMethodCoupleError) as e:
        lq.HybridMethodCouple(ov, Boom()).next_element(state)
    assert e.value.which == "ErrorSecond"


def test_metropolis_hastings_delta_diagnostic_readme(lib):
    """The README example (README.md:47-80; metropolis_hastings.rs:300-417) at a reduced size: 6^4 (README: 10^4),
    beta = 1, a = 1000, spread 0.1, single random link hits with normalize_link_matrices in between, then
    average_trace_plaquette().real() / 3.  One reference call = one hit; hits_per_call batches them on the device."""
    assert lq.MetropolisHastingsDeltaDiagnostic.new(0.0, lq.Rng(1)) is None
    assert lq.MetropolisHastingsDeltaDiagnostic.new(1.0, lq.Rng(1)) is None
    rng = lq.Rng(SEED_RNG)
    state = lq.LatticeStateDefault.new_determinist(1000.0, 1.0, 6, rng, lib=lib)
    nl = state.lattice().number_of_canonical_links_space()
    # the reference's call, one hit at a time
    mh1 = lq.MetropolisHastingsDeltaDiagnostic.new(0.1, rng)
    for _ in range(20):
        state = state.monte_carlo_step(mh1)
        assert mh1.hits_performed_last == 1 and 0.0 <= mh1.prob_replace_last() <= 1.0
        assert mh1.has_replace_last() == (mh1.hits_accepted_last == 1)
    # batched: ~40 hits per link in total
    mh = lq.MetropolisHastingsDeltaDiagnostic.new(0.1, rng, hits_per_call=nl // 16)
    vals = []
    for k in range(40 * 16):
        state = state.monte_carlo_step(mh)
        if k % 16 == 15:
            state.normalize_link_matrices()
            if k >= 20 * 16:
                vals.append(state.average_trace_plaquette().real / 3.0)
    assert 0.5 < mh.prob_replace_last() <= 1.0
    assert abs(np.mean(vals) - 1.0 / 18.0) < 0.01  # strong coupling: <P>/3 ~ beta/18


def test_hmc_rejects_host_integrators(lib):
    """a non-device integrator must not silently run whatever the context had selected."""
    class HostIntegrator:
        pass

    rng = lq.Rng(3)
    state = lq.LatticeStateDefault.new_determinist(1.0, 6.0, 4, rng, lib=lib)
    with pytest.raises(TypeError):
        lq.HybridMonteCarloDiagnostic.new(0.01, 2, HostIntegrator(), rng).next_element(state)
    # an Omelyan trajectory leaves the context on the reference's integrator
    hmc = lq.HybridMonteCarloDiagnostic.new(0.01, 2, lq.OmelyanCuda.new(), rng)
    state = hmc.next_element(state)
    ref = state.clone()
    a = lq.LatticeStateEFSyncDefault.new_random_e_state(state, lq.Rng(9))
    b = lq.LatticeStateEFSyncDefault.new_random_e_state(ref, lq.Rng(9))
    a = a.simulate_symplectic_n(lq.SymplecticEulerCuda.new(), 0.01, 2)
    b._touch().symplectic_n(0.01, 2)  # lq_symplectic_n directly: the reference's integrator
    assert np.array_equal(a.link_matrix(), b.link_matrix())
