"""The five configurations of BASELINE.json as parity cases.

Small instances compare with the oracle element-wise (they run on the CPU CI through the host-emulation backend
and on the B200 through the CUDA library); the full sizes (32^4, 40^3, 10^4) run on the GPU only and are checked
through size-independent properties: translation covariance of the stencils (an index-arithmetic check that needs
no oracle), action conservation of over-relaxation, idempotence of the SU(3) reprojection, exactness of the cold
fixed point, and statistical agreement (2 sigma) of the checkerboard sweeps with the reference's sequential order.
"""
import numpy as np
import pytest

from oracle.oracle import Oracle, to_c
from tests.conftest import SEED_RNG


@pytest.fixture(params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "emu":
        from tests import emu
        return "emu", emu.context
    import torch
    assert torch.cuda.is_available(), "the gpu-marked tests need a CUDA device (no CPU fallback)"
    from lattice_qcd_rs_b200 import Context
    return "cuda", Context


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def shift_sites(A, ext, d, per_site):
    """Field translated by one site in direction d: B(x) = A(x - e_d), reference site order (x0 fastest)."""
    D = len(ext)
    a = A.reshape(*ext[::-1], per_site)
    return np.ascontiguousarray(np.roll(a, 1, axis=D - 1 - d)).reshape(A.shape)


# ---------------------------------------------------------------------------------------------------- config 1
def test_config1_hmc_trajectories(backend):
    """LatticeStateDefault::<4> 8^4 (4^4 on the CPU CI), beta = 6, HybridMonteCarloDiagnostic + symplectic Euler,
    10 trajectories, average_trace_plaquette: identical start configuration and identical momenta streams."""
    name, ctx = backend
    n = 8 if name == "cuda" else 4
    ntraj = 10 if name == "cuda" else 3
    o = Oracle(4, n, a=1.0, beta=6.0)
    c = ctx(4, n, a=1.0, beta=6.0)
    U = o.links_random(SEED_RNG)
    c.links_upload(U)
    for k in range(ntraj):
        r = c.hmc_trajectory(0.01, 10, SEED_RNG, k)
        ro = o.hmc_trajectory(U, 0.01, 10, SEED_RNG, k, literal=False)
        assert r["accepted"] == ro["accepted"] and r["gauss_steps"] == ro["gauss_steps"]
        assert abs((r["h_new"] - r["h_old"]) - (ro["h_new"] - ro["h_old"])) <= 1e-9 * abs(ro["h_old"])
        U = ro["U"]
    assert rel(c.links_download(), U) <= 1e-9
    assert abs(c.average_trace_plaquette() - o.average_trace_plaquette(U)) <= 1e-11


# ---------------------------------------------------------------------------------------------------- config 2
def test_config2_metropolis_plaquette_agreement(backend):
    """10^4 (6^4 on the CPU CI), beta = 1, a = 1000, spread 0.1 (README.md:47-80): the reference's single-link
    Metropolis in its own order (oracle, sequential hits on random links, reprojection every 1000 hits) against
    checkerboard sweeps at matched hits per link.  <Re Tr P>/3 agrees within 2 sigma (sigma = sqrt(var/len),
    statistics/mod.rs:401-405, doubled for autocorrelation)."""
    name, ctx = backend
    n = 10 if name == "cuda" else 6
    o = Oracle(4, n, a=1000.0, beta=1.0)
    c = ctx(4, n, a=1000.0, beta=1.0)
    U0 = o.links_random(SEED_RNG)
    therm, meas = 30, 30
    # reference order: one "sweep" = Nl single-link hits on uniformly random links
    U, ref = U0, []
    for k in range(therm + meas):
        for part in range(max(o.nl // 1000, 1)):
            U, _, _ = o.metropolis_single_link(U, SEED_RNG + 1, k * 10000 + part, 0.1, min(1000, o.nl))
            U = o.normalize_links(U)
        if k >= therm:
            ref.append(o.average_trace_plaquette(U).real / 3.0)
    c.links_upload(U0)
    got, acc = [], []
    for k in range(therm + meas):
        na, sp = c.sweep_metropolis(SEED_RNG, k, spread=0.1, n_update=1)
        c.reunitarize()
        acc.append(na / o.nl)
        if k >= therm:
            got.append(c.average_trace_plaquette().real / 3.0)
    ref, got = np.array(ref), np.array(got)
    sig = np.hypot(ref.std(ddof=1) / np.sqrt(ref.size), got.std(ddof=1) / np.sqrt(got.size))
    assert abs(ref.mean() - got.mean()) < 2.0 * 2.0 * sig, (ref.mean(), got.mean(), sig)
    assert abs(got.mean() - 1.0 / 18.0) < 0.01  # strong coupling: <P>/3 ~ beta/18
    assert 0.5 < np.mean(acc) <= 1.0


# ---------------------------------------------------------------------------------------------------- config 3
@pytest.mark.gpu
def test_config3_full_size_properties():
    """32^4 beta = 6: HMC + heat-bath / over-relaxation sweeps, normalize_link_matrices -- properties that hold at
    any size.  (The oracle finishes 8^4 in seconds, not 32^4.)"""
    import torch
    assert torch.cuda.is_available()
    from lattice_qcd_rs_b200 import Context
    n = 32
    ext = [n] * 4
    c = Context(4, n, a=1.0, beta=6.0)
    # cold start: exact fixed point of the integrator (test_sim_cold), plaquette exactly 3
    c.links_set_cold()
    c.efield_set_zero()
    c.symplectic_n(0.01, 2)
    assert c.average_trace_plaquette() == 3.0 and c.hamiltonian_total() == 0.0
    # hot start: translation covariance of plaquette, force and one fused MD step (pure index arithmetic at full size)
    c.links_set_random(SEED_RNG, 0)
    c.momenta_refresh(SEED_RNG, 1)
    U, E = c.links_download(), c.efield_download()
    ps, F = c.plaquette_sum(), c.force()
    c.symplectic_n(0.01, 1)
    U1 = c.links_download()
    for d in (0, 3):
        c.links_upload(shift_sites(U, ext, d, 4 * 18))
        c.efield_upload(shift_sites(E, ext, d, 4 * 8))
        assert abs(c.plaquette_sum() - ps) <= 1e-12 * abs(ps)
        assert rel(c.force(), shift_sites(F, ext, d, 4 * 8)) <= 1e-13
        c.symplectic_n(0.01, 1)
        assert rel(c.links_download(), shift_sites(U1, ext, d, 4 * 18)) <= 1e-13
    # over-relaxation conserves the action; the heat bath moves it; reprojection is idempotent and lands in SU(3)
    c.links_upload(U)
    h = c.hamiltonian_links()
    c.sweep_overrelax(1)
    assert abs(c.hamiltonian_links() - h) <= 1e-10 * abs(h)
    for k in range(3):
        c.sweep_heatbath(SEED_RNG, 10 + k)
    assert c.hamiltonian_links() < 0.9 * h
    c.reunitarize()
    V = c.links_download()
    c.reunitarize()
    assert rel(c.links_download(), V) <= 1e-13  # max over 75 M entries of a second Gram-Schmidt pass: a few dozen ulp
    M = to_c(V[:: 4099])
    assert np.abs(M @ np.conj(np.swapaxes(M, 1, 2)) - np.eye(3)).max() <= 1e-12
    assert np.abs(np.linalg.det(M) - 1.0).max() <= 1e-12
    # one HMC trajectory at the bench's step size: the accept/reject bookkeeping is consistent
    r = c.hmc_trajectory(0.01, 20, SEED_RNG, 5)
    assert np.isfinite(r["h_new"]) and abs(r["prob"] - min(1.0, np.exp(r["h_old"] - r["h_new"]))) <= 1e-12
    assert r["gauss_steps"] % 4 == 1  # 1 + 4k steps, field.rs:1265-1294


# ---------------------------------------------------------------------------------------------------- config 4
def test_config4_anisotropic_replica(backend):
    """48^3 x 96 is not representable by LatticeCyclic (one `dim` for all directions, lattice.rs:44-49): per-direction
    extents are an extension of the C ABI, so parity is against the oracle on the reduced replica 12^3 x 24 (GPU;
    6^3 x 12 on the CPU CI).  The t-decomposed version of the same check is tests/test_dist.py."""
    name, ctx = backend
    ext = [12, 12, 12, 24] if name == "cuda" else [6, 6, 6, 12]
    o = Oracle(4, ext, a=1.0, beta=6.0)
    c = ctx(4, ext, a=1.0, beta=6.0)
    U = o.links_random(SEED_RNG)
    E = o.momenta_refresh(SEED_RNG, 9)
    c.links_upload(U)
    c.efield_upload(E)
    assert abs(c.plaquette_sum() - o.plaquette_sum(U)) <= 1e-12 * abs(o.plaquette_sum(U))
    assert rel(c.force(), o.force(U, literal=False)) <= 1e-12
    c.symplectic_n(0.01, 3)
    Uo, Eo = o.integrate(U, E, "symplectic", 0.01, n=3, literal=False)
    assert rel(c.links_download(), Uo) <= 1e-12 and rel(c.efield_download(), Eo) <= 1e-12
    c.links_upload(U)
    c.sweep_heatbath(SEED_RNG, 4)
    assert rel(c.links_download(), o.sweep_heatbath(U, SEED_RNG, 4)) <= 1e-9


# ---------------------------------------------------------------------------------------------------- config 5
def test_config5_generic_dimension(backend):
    """D = 3 40^3 (12^3 on the CPU CI) through the same dimension-generic API, beta = 6.2."""
    name, ctx = backend
    n = 40 if name == "cuda" else 12
    o = Oracle(3, n, a=1.0, beta=6.2)
    c = ctx(3, n, a=1.0, beta=6.2)
    U = o.links_random(SEED_RNG)
    E = o.momenta_refresh(SEED_RNG, 2)
    c.links_upload(U)
    c.efield_upload(E)
    assert abs(c.hamiltonian_total() - o.hamiltonian_total(U, E)) <= 1e-12 * abs(o.hamiltonian_total(U, E))
    assert rel(c.staples(), o.staples(U)) <= 1e-12
    assert rel(c.gauss_field(), o.gauss_field(U, E)) <= 1e-12
    c.symplectic_n(0.005, 4)
    Uo, Eo = o.integrate(U, E, "symplectic", 0.005, n=4, literal=False)
    assert rel(c.links_download(), Uo) <= 1e-12 and rel(c.efield_download(), Eo) <= 1e-12
    c.links_upload(U)
    for kind in (0, 1):
        h = c.hamiltonian_links()
        c.sweep_overrelax(kind)
        assert abs(c.hamiltonian_links() - h) <= 1e-10 * abs(h)
    c.sweep_heatbath(SEED_RNG, 3)
    c.reunitarize()  # the heat bath as coded leaves SU(3) (PAULI_3 quirk); HMC's Gauss projection needs unitary links
    r = c.hmc_trajectory(0.005, 5, SEED_RNG, 8)
    assert np.isfinite(r["h_new"]) and 0.0 <= r["prob"] <= 1.0


@pytest.mark.gpu
def test_config5_largest_single_gpu_lattice():
    """64^4 beta = 6.2 on ONE GPU (16.8 M sites, 9.7 GB of links: the largest lattice of BASELINE.json and close to
    the 2^31-element limit of the tuned kernels' 32-bit indices): size-independent invariants only."""
    import torch
    assert torch.cuda.is_available()
    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs ~45 GB of device memory")
    from lattice_qcd_rs_b200 import Context
    c = Context(4, 64, a=1.0, beta=6.2)
    assert c.ns == 64 ** 4 and c.nl == 4 * 64 ** 4
    c.links_set_cold()
    c.efield_set_zero()
    c.symplectic_n(0.01, 1)
    assert c.average_trace_plaquette() == 3.0 and c.hamiltonian_total() == 0.0
    c.links_set_random(SEED_RNG, 0)
    p = c.average_trace_plaquette()
    assert abs(p) < 0.01  # random SU(3): <Tr P> = 0 up to 1/sqrt(6 Ns) fluctuations
    h = c.hamiltonian_links()
    assert abs(h / (6.2 * 6 * c.ns) - 1.0) < 1e-3
    c.sweep_overrelax(1)
    assert abs(c.hamiltonian_links() - h) <= 1e-10 * abs(h)
    c.momenta_refresh(SEED_RNG, 1)
    h0 = c.hamiltonian_total()
    c.symplectic_n(1e-4, 2)
    assert abs(c.hamiltonian_total() - h0) <= 1e-6 * abs(h0) and c.t == 3
    c.reunitarize()
    c.close()


@pytest.mark.gpu
def test_config3_full_size_matches_oracle():
    """32^4 -- the size every bench number is quoted on -- CUDA against the oracle element-wise: plaquette / H_links,
    the force, one fused MD step (lq_symplectic_n(dt, 1)), the Gauss field and one projection step, one heat-bath and
    one over-relaxation sub-sweep pair, and the bulk-copy (TMA) upload / download round trip.  The tuned kernels' 32-bit
    index arithmetic and the chunk layout at ext0 = 32 (one chunk per x0 row) are exercised only here.  About a minute
    of oracle time on the box's host cores."""
    import os
    import torch
    assert torch.cuda.is_available()
    from lattice_qcd_rs_b200 import Context
    n = 32
    o = Oracle(4, n, a=1.0, beta=6.0)
    o.set_num_threads(os.cpu_count() or 1)
    c = Context(4, n, a=1.0, beta=6.0)
    U = o.links_random(SEED_RNG)
    E = o.momenta_refresh(SEED_RNG, 1)
    c.links_upload(U)
    c.efield_upload(E)
    assert np.array_equal(c.links_download(), U) and np.array_equal(c.efield_download(), E)  # TMA row kernels: bit exact
    ps = o.plaquette_sum(U)
    assert abs(c.plaquette_sum() - ps) <= 1e-12 * abs(ps)
    hl = o.hamiltonian_links(U)
    assert abs(c.hamiltonian_links() - hl) <= 1e-12 * abs(hl)
    he = o.hamiltonian_efield(E)
    assert abs(c.hamiltonian_efield() - he) <= 1e-12 * abs(he)
    assert rel(c.force(), o.force(U, literal=False)) <= 1e-12
    # Gauss law: field, residual, one projection step
    assert rel(c.gauss_field(), o.gauss_field(U, E)) <= 1e-12
    gd = o.gauss_sum_div(U, E)
    assert abs(c.gauss_sum_div() - gd) <= 1e-12 * gd
    c.gauss_project_step()
    assert rel(c.efield_download(), o.project_to_gauss_step(U, E)) <= 1e-12
    # one fused MD step (kick dt/2, link step, kick dt/2)
    c.efield_upload(E)
    c.symplectic_n(0.01, 1)
    Uo, Eo = o.integrate(U, E, "symplectic", 0.01, n=1, literal=False)
    assert rel(c.links_download(), Uo) <= 1e-12 and rel(c.efield_download(), Eo) <= 1e-12
    # local updates: a full heat-bath sweep and a full over-relaxation sweep against the oracle's checkerboard order
    c.links_upload(U)
    c.sweep_heatbath(SEED_RNG, 3)
    assert rel(c.links_download(), o.sweep_heatbath(U, SEED_RNG, 3)) <= 1e-9
    c.links_upload(U)
    c.sweep_overrelax(0)
    assert rel(c.links_download(), o.sweep_overrelax(U, 0)) <= 1e-9


# ---------------------------------------------------------------------------------------------------- config 2, README method
def test_config2_single_link_hits_agree_with_sequential_reference(backend):
    """MetropolisHastingsDeltaDiagnostic (metropolis_hastings.rs:374-417; the README example, README.md:47-80): batches
    of independent random single-link hits on the device (lq_metropolis_hits) against the reference's own sequence of
    single hits (oracle), 10^4 beta = 1 (6^4 on the CPU CI), matched hits per link, <Re Tr P>/3 within 2 sigma."""
    name, ctx = backend
    n = 10 if name == "cuda" else 6
    o = Oracle(4, n, a=1000.0, beta=1.0)
    c = ctx(4, n, a=1000.0, beta=1.0)
    U0 = o.links_random(SEED_RNG)
    therm, meas = 30, 30
    U, ref = U0, []
    for k in range(therm + meas):
        for part in range(max(o.nl // 1000, 1)):
            U, _, _ = o.metropolis_single_link(U, SEED_RNG + 1, k * 10000 + part, 0.1, min(1000, o.nl))
            U = o.normalize_links(U)
        if k >= therm:
            ref.append(o.average_trace_plaquette(U).real / 3.0)
    c.links_upload(U0)
    got, perf_frac, acc = [], [], []
    batch = o.nl // 16  # hits per call: an eighth of the links are in the call's (direction, colour) class
    for k in range(therm + meas):
        done = 0
        part = 0
        while done < o.nl:  # one "sweep" = Nl performed hits
            n_perf, n_acc, sum_p = c.metropolis_hits(SEED_RNG + 2, k * 100000 + part, 0.1, batch)
            assert 0 < n_perf <= batch and n_acc <= n_perf and 0.0 <= sum_p <= n_perf
            perf_frac.append(n_perf / batch)
            acc.append(n_acc / n_perf)
            done += n_perf
            part += 1
        c.reunitarize()
        if k >= therm:
            got.append(c.average_trace_plaquette().real / 3.0)
    ref, got = np.array(ref), np.array(got)
    sig = np.hypot(ref.std(ddof=1) / np.sqrt(ref.size), got.std(ddof=1) / np.sqrt(got.size))
    assert abs(ref.mean() - got.mean()) < 2.0 * 2.0 * sig, (ref.mean(), got.mean(), sig)
    assert abs(got.mean() - 1.0 / 18.0) < 0.01
    assert 0.5 < np.mean(acc) <= 1.0
    # collisions inside a batch: n hits on m = Nl/8 links leave m (1 - exp(-n/m)) distinct ones: 1 - e^-0.5 = 0.787 of n
    assert abs(np.mean(perf_frac) - (1.0 - np.exp(-0.5)) / 0.5) < 0.02


def test_single_link_hit_is_the_reference_call(backend):
    """n_hits = 1: one uniformly random link, same proposal and accept rule as delta_s_old_new_cmp; exactly one link
    changes when the hit is accepted and none otherwise; force_accept applies the proposal unconditionally."""
    name, ctx = backend
    o = Oracle(4, 4, a=1.0, beta=6.0)
    c = ctx(4, 4, a=1.0, beta=6.0)
    U = o.links_random(SEED_RNG)
    c.links_upload(U)
    changed = 0
    dirs = set()
    for k in range(40):
        before = c.links_download().copy()
        n_perf, n_acc, sum_p = c.metropolis_hits(SEED_RNG, 1000 + k, 0.3, 1)
        after = c.links_download()
        diff = np.flatnonzero(np.abs(after - before).max(axis=1) > 0)
        assert n_perf == 1 and n_acc in (0, 1) and 0.0 <= sum_p <= 1.0
        assert diff.size == n_acc
        if n_acc:
            changed += 1
            dirs.add(int(diff[0]) % 4)
            # the accepted matrix is (an SU(3) matrix close to one) x (the old link): still unitary to rounding
            m = to_c(after[diff])
            assert np.abs(m @ np.conj(np.swapaxes(m, 1, 2)) - np.eye(3)).max() < 1e-12
    assert 0 < changed < 40 and len(dirs) >= 3
    before = c.links_download().copy()
    n_perf, n_acc, _ = c.metropolis_hits(SEED_RNG, 5000, 0.3, 64, force_accept=True)
    diff = np.flatnonzero(np.abs(c.links_download() - before).max(axis=1) > 0)
    assert n_acc == n_perf == diff.size and 32 <= n_perf <= 64
    assert len(set(int(i) % 4 for i in diff)) == 1  # one (direction, colour) class per call
