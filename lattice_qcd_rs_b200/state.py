"""Host-side mirror of the lattice-qcd-rs v0.2.1 trait surface for the pure-gauge update path.

The reference is a Rust crate; this image has no rustc, so the host side above the C ABI is written here with the
reference's own names, argument meaning and error behaviour (the authored Rust shim with the same shape lives in
rust/lattice-qcd-b200/).  Everything below is bookkeeping around one `lq_ctx` per state: no
lattice arithmetic happens in Python, and there is no CPU fallback (creating a state without the CUDA library or a
device raises).

  reference item (file:line under /root/reference/src)                      here
  ------------------------------------------------------------------------  ---------------------------------------
  LatticeCyclic<D>::new(size, dim)                 lattice.rs:190-201        LatticeCyclic
  LatticeStateDefault<D> + LatticeState/New        state.rs:52-154,655-849   LatticeStateDefault
  LatticeStateEFSyncDefault + ...WithEField/New    state.rs:165-278,1048-    LatticeStateEFSyncDefault
  SimulationStateSynchronous / ...LeapFrog         state.rs:292-648          methods simulate_* of the two states
  SimulationStateLeap                              state.rs:856-861          SimulationStateLeap
  SymplecticIntegrator / SymplecticEulerRayon      integrator/mod.rs:93-208  SymplecticEulerCuda
  MonteCarlo::next_element                         monte_carlo/mod.rs:65-77  every method class below
  HybridMonteCarlo(Diagnostic)                     hybrid_monte_carlo.rs     HybridMonteCarlo, ...Diagnostic
  HeatBathSweep                                    heat_bath.rs:40-157       HeatBathSweep
  OverrelaxationSweepRotation / Reverse            overrelaxation.rs         OverrelaxationSweepRotation / Reverse
  MetropolisHastingsSweep                          metropolis_hastings_sweep MetropolisHastingsSweep
  HybridMethodVec / HybridMethodCouple(/Triple..)  hybrid.rs:248-268,330-446 HybridMethodVec, HybridMethodCouple, ...
  MonteCarloDefault + McWrapper                    monte_carlo/mod.rs:117-293 MonteCarloDefault, McWrapper
  MetropolisHastings(Diagnostic)                   metropolis_hastings.rs:40-  MetropolisHastings, ...Diagnostic
  MetropolisHastingsDeltaDiagnostic                metropolis_hastings.rs:300- MetropolisHastingsDeltaDiagnostic
  StateInitializationError / MultiIntegrationError error.rs:93-133           exceptions of the same names
  serde derives (feature serde-serialize)          state.rs:654, 1047        to_json / to_bincode / from_* (serde_io.py)
  -- not in the crate, same surfaces -------------------------------------------------------------
  Omelyan steps, exponential link update           (integrator/mod.rs:93)    OmelyanCuda (a SymplecticIntegrator)
  SU(2)-sub-group over-relaxation                  (overrelaxation.rs)       OverrelaxationSweepSu2

Differences that the drop-in cannot hide: sweeps visit links in even/odd checkerboard order
instead of the reference's sequential index order, and every stochastic draw comes from Philox4x32-10 streams keyed
by (seed, call counter, global link index) -- the host `rng` argument is only asked for one u64 per call.
"""
import math

import numpy as np

from . import _capi, serde_io
from ._capi import Context, LqError

CA = 3.0  # LatticeState::CA, state.rs:796


# ------------------------------------------------------------------------------------------------ errors (error.rs)
class LatticeInitializationError(ValueError):
    """error.rs:219-226: NonPositiveSize | DimTooSmall | ZeroDimension"""


class StateInitializationError(ValueError):
    """error.rs:124-133: InvalidParameterNormal | IncompatibleSize | LatticeInitializationError | GaussProjectionError"""

    def __init__(self, kind, detail=""):
        self.kind = kind
        super().__init__(f"{kind} {detail}".strip())


class MultiIntegrationError(RuntimeError):
    """error.rs:93-98: ZeroIntegration | IntegrationError(step, error)"""

    def __init__(self, kind, step=None, error=None):
        self.kind, self.step, self.error = kind, step, error
        super().__init__(kind if step is None else f"{kind}({step}, {error})")


def _wrap(err):
    """C error code -> the reference's error enum."""
    if isinstance(err, LqError):
        if err.code == -2:
            return StateInitializationError("IncompatibleSize")
        if err.code == -6:
            return StateInitializationError("GaussProjectionError")
        if err.code == -7:
            return MultiIntegrationError("ZeroIntegration")
    return err


# ------------------------------------------------------------------------------------------------ host rng
class Rng:
    """Stand-in for the `rand::Rng` argument of the reference API (e.g. StdRng::seed_from_u64, test/mod.rs:19).
    The device draws from Philox streams; the host generator is only asked for one u64 per Monte-Carlo call
    (SplitMix64), so a run is reproducible from the seed alone."""

    def __init__(self, seed=0x457893F44AB067F0):
        self.state = int(seed) & 0xFFFFFFFFFFFFFFFF

    @classmethod
    def seed_from_u64(cls, seed):
        return cls(seed)

    def checkpoint(self):
        """The whole generator state (one u64): together with a serialized lattice state this resumes a run bit for
        bit, because every device draw is a Philox stream keyed by the (seed, counter) pair drawn here."""
        return self.state

    @classmethod
    def from_checkpoint(cls, state):
        r = cls(0)
        r.state = int(state) & 0xFFFFFFFFFFFFFFFF
        return r

    def next_u64(self):
        self.state = (self.state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)


def _draw(rng):
    """(seed, counter) of one device call: a fresh 64-bit seed and a 56-bit counter from the host generator."""
    return rng.next_u64(), rng.next_u64() >> 8


# ------------------------------------------------------------------------------------------------ lattice
class LatticeCyclic:
    """lattice.rs:44-49, 190-201: hyper-cubic periodic lattice, `dim` points per direction, physical `size`."""

    def __init__(self, size, dim, D=4):
        if D == 0:
            raise LatticeInitializationError("ZeroDimension")
        if not (size > 0.0) or math.isnan(size) or math.isinf(size):
            raise LatticeInitializationError("NonPositiveSize")
        if dim < 2:
            raise LatticeInitializationError("DimTooSmall")
        self._size, self._dim, self.D = float(size), int(dim), int(D)

    @classmethod
    def new(cls, size, dim, D=4):
        return cls(size, dim, D)

    def size(self):
        return self._size

    def dim(self):
        return self._dim

    def number_of_points(self):            # lattice.rs get_number_of_points
        return self._dim ** self.D

    def number_of_canonical_links_space(self):
        return self.number_of_points() * self.D

    def has_compatible_length_links(self, link_matrix):
        return np.asarray(link_matrix).size == self.number_of_canonical_links_space() * 18

    def has_compatible_length_e_field(self, e_field):
        return np.asarray(e_field).size == self.number_of_canonical_links_space() * 8

    def __eq__(self, o):
        return isinstance(o, LatticeCyclic) and (self._size, self._dim, self.D) == (o._size, o._dim, o.D)


# ------------------------------------------------------------------------------------------------ LatticeState
class LatticeStateDefault:
    """state.rs:655-849.  Links live on the device; `link_matrix()` returns a lazily refreshed host mirror in the
    reference AoS layout ((Nl, 18) f64: link = site*D + dir, 3x3 complex column-major)."""

    CA = CA

    def __init__(self, ctx, lattice):
        self._ctx, self._lattice = ctx, lattice
        self._host_links = None

    # -- constructors -------------------------------------------------------------------------------------------
    @staticmethod
    def _ctx_for(lattice, beta, device, lib):
        try:
            return Context(lattice.D, lattice.dim(), a=lattice.size(), beta=beta, CA=CA, device=device, lib=lib)
        except LqError as e:
            raise _wrap(e)

    @classmethod
    def new(cls, lattice, beta, link_matrix, device=0, lib=None):
        """LatticeStateNew::new, state.rs:779-792."""
        if not lattice.has_compatible_length_links(link_matrix):
            raise StateInitializationError("IncompatibleSize")
        st = cls(cls._ctx_for(lattice, beta, device, lib), lattice)
        st.set_link_matrix(link_matrix)
        return st

    @classmethod
    def new_cold(cls, size, beta, number_of_points, D=4, device=0, lib=None):
        """state.rs:671-679."""
        try:
            lattice = LatticeCyclic(size, number_of_points, D)
        except LatticeInitializationError as e:
            raise StateInitializationError("LatticeInitializationError", str(e))
        st = cls(cls._ctx_for(lattice, beta, device, lib), lattice)
        st._ctx.links_set_cold()
        return st

    @classmethod
    def new_determinist(cls, size, beta, number_of_points, rng, D=4, device=0, lib=None):
        """state.rs:706-715: random_su3 per link (su3.rs:322-355), drawn on the device."""
        try:
            lattice = LatticeCyclic(size, number_of_points, D)
        except LatticeInitializationError as e:
            raise StateInitializationError("LatticeInitializationError", str(e))
        st = cls(cls._ctx_for(lattice, beta, device, lib), lattice)
        st._ctx.links_set_random(*_draw(rng))
        return st

    # -- LatticeState ----------------------------------------------------------------------------------------------
    def link_matrix(self):
        if self._host_links is None:
            self._host_links = self._ctx.links_download()
            self._host_links.setflags(write=False)
        return self._host_links

    def set_link_matrix(self, link_matrix):
        """Panics (AssertionError) on a wrong length, state.rs:808-815."""
        assert self._lattice.has_compatible_length_links(link_matrix), "Link matrices are not of the correct size"
        self._ctx.links_upload(np.asarray(link_matrix, dtype=np.float64).reshape(-1, 18))
        self._host_links = None

    def lattice(self):
        return self._lattice

    def beta(self):
        return self._ctx.beta

    def hamiltonian_links(self):
        return self._ctx.hamiltonian_links()

    def average_trace_plaquette(self):
        return self._ctx.average_trace_plaquette()

    def monte_carlo_step(self, m):
        """state.rs:103-109: consumes self, returns the next state."""
        return m.next_element(self)

    def normalize_link_matrices(self):
        """state.rs:754-756."""
        self._ctx.reunitarize()
        self._host_links = None

    def link_matrix_owned(self):
        return np.array(self.link_matrix())

    def set_flags(self, flags):
        """Behaviour switches of the device state (include/lqcd_b200.h LQ_FLAG_*): 0 restates the crate as coded;
        FLAG_PAULI3_FIXED | FLAG_UNIFORM_DIRECTION give the textbook heat bath."""
        self._ctx.set_flags(flags)
        return self

    def clone(self):
        st = type(self)(self._ctx.clone(), self._lattice)
        return st

    # -- serde (feature `serde-serialize`, state.rs:654): what serde_json / bincode 1.x write for this struct; the
    #    links travel device -> host AoS -> bytes, the layouts are documented in serde_io.py
    def to_json(self):
        lat = self._lattice
        return serde_io.dumps_json(serde_io.state_to_json_obj(lat.size(), lat.dim(), lat.D, self.beta(), self.link_matrix()))

    def to_bincode(self, seq_prefix=True):
        lat = self._lattice
        return serde_io.state_to_bincode(lat.size(), lat.dim(), lat.D, self.beta(), self.link_matrix(), seq_prefix)

    @classmethod
    def _from_plain(cls, d, device, lib):
        try:
            lattice = LatticeCyclic(d["size"], d["dim"], d["D"])
        except LatticeInitializationError as e:
            raise StateInitializationError("LatticeInitializationError", str(e))
        return cls.new(lattice, d["beta"], d["links"], device=device, lib=lib)

    @classmethod
    def from_json(cls, text, D=4, device=0, lib=None):
        import json
        return cls._from_plain(serde_io.state_from_json_obj(json.loads(text), D), device, lib)

    @classmethod
    def from_bincode(cls, buf, D=4, device=0, lib=None, seq_prefix=True):
        return cls._from_plain(serde_io.state_from_bincode(buf, D, seq_prefix), device, lib)

    # -- device handle for the method classes
    def _touch(self):
        self._host_links = None
        return self._ctx


class LatticeStateEFSyncDefault(LatticeStateDefault):
    """state.rs:1048-1062 (+ LatticeStateWithEField :165-232, SimulationStateSynchronous :292-511): links, E-field
    and the step counter t, all on the device."""

    def __init__(self, ctx, lattice):
        super().__init__(ctx, lattice)
        self._host_e = None

    @classmethod
    def new(cls, lattice, beta, e_field, link_matrix, t, device=0, lib=None):
        """LatticeStateWithEFieldNew::new, state.rs:1342-1361."""
        if not lattice.has_compatible_length_links(link_matrix) or not lattice.has_compatible_length_e_field(e_field):
            raise StateInitializationError("IncompatibleSize")
        st = cls(cls._ctx_for(lattice, beta, device, lib), lattice)
        st.set_link_matrix(link_matrix)
        st.set_e_field(e_field)
        st._ctx.set_t(t)
        return st

    @classmethod
    def new_random_e_state(cls, lattice_state, rng):
        """state.rs:1093-1108: Normal(0, 0.5/beta) momenta, Gauss-projected; takes ownership of `lattice_state`."""
        st = cls(lattice_state._ctx, lattice_state._lattice)
        lattice_state._ctx = None
        st.reset_e_field(rng)
        return st

    @classmethod
    def new_e_cold(cls, lattice_state):
        """state.rs:1111-1121."""
        st = cls(lattice_state._ctx, lattice_state._lattice)
        lattice_state._ctx = None
        st._ctx.efield_set_zero()
        return st

    @classmethod
    def new_random_e(cls, lattice, beta, link_matrix, rng, device=0, lib=None):
        """LatticeStateWithEFieldNew::new_random_e, state.rs:262-277."""
        return cls.new_random_e_state(LatticeStateDefault.new(lattice, beta, link_matrix, device, lib), rng)

    @classmethod
    def new_cold(cls, size, beta, number_of_points, D=4, device=0, lib=None):
        """state.rs:1221-1230."""
        return cls.new_e_cold(LatticeStateDefault.new_cold(size, beta, number_of_points, D, device, lib))

    @classmethod
    def new_determinist(cls, size, beta, number_of_points, rng, D=4, device=0, lib=None):
        """state.rs:1175-1194: hot links, random Gauss-projected E."""
        return cls.new_random_e_state(
            LatticeStateDefault.new_determinist(size, beta, number_of_points, rng, D, device, lib), rng)

    # -- LatticeStateWithEField ------------------------------------------------------------------------------------
    def reset_e_field(self, rng):
        """state.rs:174-189."""
        if not (self.beta() != 0.0 and math.isfinite(0.5 / self.beta())):
            raise StateInitializationError("InvalidParameterNormal")
        seed, counter = _draw(rng)
        try:
            self._ctx.momenta_refresh(seed, counter, 0.5 / self.beta())
            self._ctx.gauss_project()
        except LqError as e:
            raise _wrap(e)
        self._host_e = None

    def e_field(self):
        if self._host_e is None:
            self._host_e = self._ctx.efield_download()
            self._host_e.setflags(write=False)
        return self._host_e

    def set_e_field(self, e_field):
        assert self._lattice.has_compatible_length_e_field(e_field), "E field is not of the correct size"
        self._ctx.efield_upload(np.asarray(e_field, dtype=np.float64).reshape(-1, 8))
        self._host_e = None

    def t(self):
        return self._ctx.t

    def hamiltonian_efield(self):
        return self._ctx.hamiltonian_efield()

    def hamiltonian_total(self):
        return self._ctx.hamiltonian_total()

    def gauss(self):
        """EField::gauss for every site (field.rs:1174-1195), (Ns, 18) AoS."""
        return self._ctx.gauss_field()

    def lattice_state(self):
        return self

    def state_owned(self):
        """state.rs:1071-1076: drop E, keep the links (same device buffers)."""
        st = LatticeStateDefault(self._ctx, self._lattice)
        self._ctx = None
        return st

    # -- serde (state.rs:1047-1062: e_field, t, lattice_state)
    def to_json(self):
        lat = self._lattice
        return serde_io.dumps_json(serde_io.ef_state_to_json_obj(lat.size(), lat.dim(), lat.D, self.beta(),
                                                                 self.link_matrix(), self.e_field(), self.t()))

    def to_bincode(self, seq_prefix=True):
        lat = self._lattice
        return serde_io.ef_state_to_bincode(lat.size(), lat.dim(), lat.D, self.beta(), self.link_matrix(),
                                            self.e_field(), self.t(), seq_prefix)

    @classmethod
    def _from_plain(cls, d, device, lib):
        try:
            lattice = LatticeCyclic(d["size"], d["dim"], d["D"])
        except LatticeInitializationError as e:
            raise StateInitializationError("LatticeInitializationError", str(e))
        return cls.new(lattice, d["beta"], d["e_field"], d["links"], d["t"], device=device, lib=lib)

    @classmethod
    def from_json(cls, text, D=4, device=0, lib=None):
        import json
        return cls._from_plain(serde_io.ef_state_from_json_obj(json.loads(text), D), device, lib)

    @classmethod
    def from_bincode(cls, buf, D=4, device=0, lib=None, seq_prefix=True):
        return cls._from_plain(serde_io.ef_state_from_bincode(buf, D, seq_prefix), device, lib)

    def _touch(self):
        self._host_e = None
        return super()._touch()

    # -- SimulationStateSynchronous (state.rs:292-511) -------------------------------------------------------------
    def simulate_sync(self, integrator, delta_t):
        return integrator.integrate_sync_sync(self, delta_t)

    def simulate_sync_n(self, integrator, delta_t, numbers_of_times):
        return _n_times(self, numbers_of_times, lambda s: integrator.integrate_sync_sync(s, delta_t))

    def simulate_symplectic(self, integrator, delta_t):
        return integrator.integrate_symplectic(self, delta_t)

    def simulate_symplectic_n(self, integrator, delta_t, numbers_of_times):
        """state.rs:470-492; one fused device loop when the integrator is the CUDA one."""
        if numbers_of_times == 0:
            raise MultiIntegrationError("ZeroIntegration")
        if isinstance(integrator, SymplecticEulerCuda):
            new = self.clone()
            try:
                ctx = new._touch()
                integrator._select(ctx)  # the reference's symplectic Euler (lq_symplectic_n) unless it is OmelyanCuda
                ctx.md_n(delta_t, numbers_of_times)
                SymplecticEulerCuda._select(integrator, ctx)
            except LqError as e:
                raise _wrap(e)
            return new
        return _n_times(self, numbers_of_times, lambda s: integrator.integrate_symplectic(s, delta_t))

    def simulate_symplectic_n_auto(self, integrator, delta_t, number_of_steps):
        return self.simulate_symplectic_n(integrator, delta_t, number_of_steps)

    def simulate_to_leapfrog(self, integrator, delta_t):
        return integrator.integrate_sync_leap(self, delta_t)

    def simulate_using_leapfrog_n(self, integrator, delta_t, numbers_of_times):
        """state.rs:321-358: sync->leap, (n-1) x leap->leap, leap->sync."""
        if numbers_of_times == 0:
            raise MultiIntegrationError("ZeroIntegration")
        try:
            leap = self.simulate_to_leapfrog(integrator, delta_t)
        except Exception as e:
            raise MultiIntegrationError("IntegrationError", 0, e)
        if numbers_of_times > 1:
            try:
                leap = leap.simulate_leap_n(integrator, delta_t, numbers_of_times - 1)
            except MultiIntegrationError as e:
                if e.kind == "IntegrationError":
                    raise MultiIntegrationError("IntegrationError", e.step + 1, e.error)
                raise
        try:
            return leap.simulate_to_synchronous(integrator, delta_t)
        except Exception as e:
            raise MultiIntegrationError("IntegrationError", numbers_of_times, e)

    def simulate_using_leapfrog_n_auto(self, integrator, delta_t, number_of_steps):
        return self.simulate_using_leapfrog_n(integrator, delta_t, number_of_steps)


def _n_times(state, n, step):
    """The n-step loops of state.rs:397-419 / 470-492 / 624-647 with their error bookkeeping."""
    if n == 0:
        raise MultiIntegrationError("ZeroIntegration")
    for k in range(n):
        try:
            state = step(state)
        except Exception as e:
            raise MultiIntegrationError("IntegrationError", k, e)
    return state


class SimulationStateLeap:
    """state.rs:856-861: a synchronous state whose E-field is half a step ahead."""

    def __init__(self, state):
        self._state = state

    @classmethod
    def new_from_state(cls, state):
        return cls(state)

    @classmethod
    def from_synchronous(cls, s, integrator, delta_t):
        return s.simulate_to_leapfrog(integrator, delta_t)

    def as_ref(self):
        return self._state

    def __getattr__(self, name):  # LatticeState / LatticeStateWithEField delegate to the inner state (state.rs:945-1045)
        return getattr(self._state, name)

    def simulate_to_synchronous(self, integrator, delta_t):
        return integrator.integrate_leap_sync(self, delta_t)

    def simulate_leap(self, integrator, delta_t):
        return integrator.integrate_leap_leap(self, delta_t)

    def simulate_leap_n(self, integrator, delta_t, numbers_of_times):
        return _n_times(self, numbers_of_times, lambda s: integrator.integrate_leap_leap(s, delta_t))


# ------------------------------------------------------------------------------------------------ integrator
class SymplecticEulerCuda:
    """SymplecticIntegrator (integrator/mod.rs:93-208) with the arithmetic of SymplecticEulerRayon
    (symplectic_euler_rayon.rs:120-252) on the device.  `&self` methods return NEW states, as in the reference."""

    @classmethod
    def new(cls):
        return cls()

    @staticmethod
    def _step(state, kind, delta_t):
        new = state.clone()
        try:
            new._touch().integrate(kind, delta_t)
        except LqError as e:
            raise _wrap(e)
        return new

    def integrate_sync_sync(self, l, delta_t):
        return self._step(l, _capi.SYNC_SYNC, delta_t)

    def integrate_leap_leap(self, l, delta_t):
        return SimulationStateLeap(self._step(l.as_ref(), _capi.LEAP_LEAP, delta_t))

    def integrate_sync_leap(self, l, delta_t):
        return SimulationStateLeap(self._step(l, _capi.SYNC_LEAP, delta_t))

    def integrate_leap_sync(self, l, delta_t):
        return self._step(l.as_ref(), _capi.LEAP_SYNC, delta_t)

    def integrate_symplectic(self, l, delta_t):
        return self._step(l, _capi.SYMPLECTIC, delta_t)

    def _select(self, ctx):
        ctx.set_integrator(_capi.INTEGRATOR_SYMPLECTIC_EULER, _capi.OMELYAN_LAMBDA, False)


class OmelyanCuda(SymplecticEulerCuda):
    """An integrator the crate does not have, offered through the same SymplecticIntegrator
    surface: `integrate_symplectic` is one second-order minimum-norm step
        E(l dt) U(dt/2) E((1-2l) dt) U(dt/2) E(l dt),   l = 0.1931833275037836,
    built from the reference's own two updates (integrate_efield, integrator/mod.rs:240-254; integrate_link,
    :216-233), with `use_exp` replacing the Euler link update by U <- exp(i dt E) U (su3.rs:832-855: links stay in
    SU(3), the step is time-reversible).  The leap-frog half-step compositions are inherited unchanged.  Passing it to
    HybridMonteCarlo(Diagnostic) makes the trajectory use it."""

    def __init__(self, lam=_capi.OMELYAN_LAMBDA, use_exp=True):
        self.lam, self.use_exp = float(lam), bool(use_exp)

    @classmethod
    def new(cls, lam=_capi.OMELYAN_LAMBDA, use_exp=True):
        return cls(lam, use_exp)

    def _select(self, ctx):
        ctx.set_integrator(_capi.INTEGRATOR_OMELYAN, self.lam, self.use_exp)

    def integrate_symplectic(self, l, delta_t):
        new = l.clone()
        try:
            ctx = new._touch()
            self._select(ctx)
            ctx.md_n(delta_t, 1)
            SymplecticEulerCuda._select(self, ctx)
        except LqError as e:
            raise _wrap(e)
        return new


# ------------------------------------------------------------------------------------------------ Monte-Carlo methods
class MonteCarlo:
    """monte_carlo/mod.rs:65-77: next_element(state) consumes the state and returns the next one."""

    def next_element(self, state):
        raise NotImplementedError


class HybridMonteCarloDiagnostic(MonteCarlo):
    """hybrid_monte_carlo.rs:316-471, 573-613: refresh momenta (sigma = 0.5/beta) + Gauss projection, n symplectic
    steps, accept with probability clamp(exp(H_old - H_new), 0, 1); the old links are kept on the device for the
    reject path."""

    def __init__(self, delta_t, number_of_steps, integrator, rng):
        self._dt, self._n, self._integrator, self._rng = float(delta_t), int(number_of_steps), integrator, rng
        self._prob_replace_last, self._has_replace_last = 0.0, False
        self.gauss_steps_last = 0

    new = classmethod(lambda cls, delta_t, number_of_steps, integrator, rng: cls(delta_t, number_of_steps, integrator,
                                                                              rng))

    def delta_t(self):
        return self._dt

    def number_of_steps(self):
        return self._n

    def integrator(self):
        return self._integrator

    def rng(self):
        return self._rng

    def rng_mut(self):
        return self._rng

    def rng_owned(self):
        return self._rng

    def prob_replace_last(self):
        return self._prob_replace_last

    def has_replace_last(self):
        return self._has_replace_last

    def next_element(self, state):
        if self._n == 0:
            raise MultiIntegrationError("ZeroIntegration")
        if not isinstance(self._integrator, SymplecticEulerCuda):
            # the trajectory runs on the device: only the device integrators can drive it (the reference runs whatever
            # SymplecticIntegrator it is given; a host integrator here would silently be replaced by another one)
            raise TypeError("HybridMonteCarlo on a device state needs a SymplecticEulerCuda or OmelyanCuda integrator, "
                            f"not {type(self._integrator).__name__}")
        if not isinstance(state, LatticeStateDefault):
            raise TypeError(f"expected a LatticeStateDefault (device state), not {type(state).__name__}")
        seed, counter = _draw(self._rng)
        try:
            ctx = state._touch()
            self._integrator._select(ctx)  # SymplecticEulerCuda: the reference's integrator; OmelyanCuda: the option
            try:
                r = ctx.hmc_trajectory(self._dt, self._n, seed, counter, sigma=0.5 / state.beta())
            finally:
                SymplecticEulerCuda._select(self._integrator, ctx)  # leave the context on the reference's integrator
        except LqError as e:
            raise _wrap(e)
        self._prob_replace_last, self._has_replace_last = r["prob"], r["accepted"]
        self.gauss_steps_last = r["gauss_steps"]
        return state


class HybridMonteCarlo(HybridMonteCarloDiagnostic):
    """hybrid_monte_carlo.rs:64-77: same algorithm without the public diagnostics."""


class HeatBathSweep(MonteCarlo):
    """heat_bath.rs:40-157: Cabibbo-Marinari r, s, t sub-group heat bath on every link (checkerboard order).
    `coupling_scale` = 1 restates the reference (Kennedy-Pendleton parameter beta*k, heat_bath.rs:77)."""

    def __init__(self, rng, coupling_scale=1.0):
        self._rng, self.coupling_scale = rng, coupling_scale

    new = classmethod(lambda cls, rng: cls(rng))

    def rng(self):
        return self._rng

    def rng_owned(self):
        return self._rng

    def next_element(self, state):
        seed, counter = _draw(self._rng)
        try:
            state._touch().sweep_heatbath(seed, counter, self.coupling_scale)
        except LqError as e:
            raise _wrap(e)
        return state


class _Overrelax(MonteCarlo):
    KIND = None

    @classmethod
    def new(cls):
        return cls()

    def next_element(self, state):
        try:
            state._touch().sweep_overrelax(self.KIND)
        except LqError as e:
            raise _wrap(e)
        return state


class OverrelaxationSweepRotation(_Overrelax):
    """overrelaxation.rs:58-110."""
    KIND = _capi.OR_ROTATION


class OverrelaxationSweepReverse(_Overrelax):
    """overrelaxation.rs:130-184."""
    KIND = _capi.OR_REVERSE


class OverrelaxationSweepSu2(_Overrelax):
    """Not in the crate: Brown-Woch reflections in the three SU(2) sub-groups of the heat bath; unlike the
    two SVD variants above (U(3)-valued, overrelaxation.rs:96-97) the links stay in SU(3)."""
    KIND = _capi.OR_SU2_SUBGROUPS


class MetropolisHastingsSweep(MonteCarlo):
    """metropolis_hastings_sweep.rs:41-174."""

    def __init__(self, number_of_update, spread, rng):
        self._n, self._spread, self._rng = int(number_of_update), float(spread), rng
        self._number_replace_last, self._prob_replace_mean = 0, 0.0

    @classmethod
    def new(cls, number_of_update, spread, rng):
        """Returns None for invalid parameters (metropolis_hastings_sweep.rs:73-80)."""
        if number_of_update == 0 or spread <= 0.0 or spread >= 1.0:
            return None
        return cls(number_of_update, spread, rng)

    def prob_replace_mean(self):
        return self._prob_replace_mean

    def number_replace_last(self):
        return self._number_replace_last

    def rng(self):
        return self._rng

    def rng_owned(self):
        return self._rng

    def next_element(self, state):
        seed, counter = _draw(self._rng)
        ctx = state._touch()
        try:
            n_acc, sum_p = ctx.sweep_metropolis(seed, counter, self._spread, self._n)
        except LqError as e:
            raise _wrap(e)
        self._number_replace_last = n_acc
        self._prob_replace_mean = sum_p / state.lattice().number_of_canonical_links_space()  # :150-172
        return state


class HybridMethodVec(MonteCarlo):
    """hybrid.rs:248-268: apply the methods one after the other."""

    def __init__(self, methods=None):
        self._methods = list(methods or [])

    def push_method(self, m):
        self._methods.append(m)

    def methods(self):
        return self._methods

    def next_element(self, state):
        for m in self._methods:
            state = m.next_element(state)
        return state


class HybridMethodCoupleError(RuntimeError):
    """hybrid.rs:281-293: ErrorFirst(e) | ErrorSecond(e)."""

    def __init__(self, which, error):
        self.which, self.error = which, error
        super().__init__(f"{which}({error!r})")


class HybridMethodCouple(MonteCarlo):
    """hybrid.rs:330-393: two methods, one after the other; errors are tagged with the method that raised them."""

    def __init__(self, method_1, method_2):
        self._m1, self._m2 = method_1, method_2

    new = classmethod(lambda cls, method_1, method_2: cls(method_1, method_2))

    def method_1(self):
        return self._m1

    def method_2(self):
        return self._m2

    def deconstruct(self):
        return self._m1, self._m2

    def next_element(self, state):
        try:
            state = state.monte_carlo_step(self._m1)
        except Exception as e:  # noqa: BLE001 -- the reference maps every error of method 1
            raise HybridMethodCoupleError("ErrorFirst", e)
        try:
            return state.monte_carlo_step(self._m2)
        except Exception as e:  # noqa: BLE001
            raise HybridMethodCoupleError("ErrorSecond", e)


def HybridMethodTriple(method_1, method_2, method_3):
    """hybrid.rs:396-404: Couple(Couple(1, 2), 3)."""
    return HybridMethodCouple(HybridMethodCouple(method_1, method_2), method_3)


def HybridMethodQuadruple(method_1, method_2, method_3, method_4):
    """hybrid.rs:410-429."""
    return HybridMethodCouple(HybridMethodTriple(method_1, method_2, method_3), method_4)


def HybridMethodQuintuple(method_1, method_2, method_3, method_4, method_5):
    """hybrid.rs:435-446."""
    return HybridMethodCouple(HybridMethodQuadruple(method_1, method_2, method_3, method_4), method_5)


class MonteCarloDefault:
    """monte_carlo/mod.rs:117-170: a method that only PROPOSES a state; `next_element_default` accepts it with
    probability `probability_of_replacement(old, new)` = clamp(exp(H_links(old) - H_links(new)), 0, 1) -- both energies
    are device reductions -- drawing the Bernoulli from the host generator, as the reference does from its `rng`."""

    def potential_next_element(self, state, rng):
        raise NotImplementedError

    @staticmethod
    def probability_of_replacement(old_state, new_state):
        d = old_state.hamiltonian_links() - new_state.hamiltonian_links()
        return max(min(math.exp(d) if d < 700.0 else math.inf, 1.0), 0.0)

    def next_element_default(self, state, rng):
        potential_next = self.potential_next_element(state, rng)
        proba = max(min(self.probability_of_replacement(state, potential_next), 1.0), 0.0)
        if _bernoulli(rng, proba):
            return potential_next
        return state


def _bernoulli(rng, p):
    """rand::distributions::Bernoulli from one u64 of the host generator."""
    return (rng.next_u64() >> 11) * (1.0 / 9007199254740992.0) < p


class McWrapper(MonteCarlo):
    """monte_carlo/mod.rs:210-293: makes a MonteCarlo out of a MonteCarloDefault and a generator."""

    def __init__(self, mcd, rng):
        self._mcd, self._rng = mcd, rng

    new = classmethod(lambda cls, mcd, rng: cls(mcd, rng))

    def mcd(self):
        return self._mcd

    def rng(self):
        return self._rng

    def rng_mut(self):
        return self._rng

    def deconstruct(self):
        return self._mcd, self._rng

    def next_element(self, state):
        return self._mcd.next_element_default(state, self._rng)


class MetropolisHastings(MonteCarloDefault):
    """metropolis_hastings.rs:40-118: proposes a state in which `number_of_update` uniformly random links were multiplied
    by orthonormalize(random_su3_close_to_unity(spread)); the whole state is then accepted or rejected on H_links
    (MonteCarloDefault).  The proposal is a device clone + one batch of forced single-link hits (`lq_metropolis_hits`,
    force_accept); hits that landed on a link already hit in the batch are dropped, so slightly fewer than
    `number_of_update` links may change.  The reference calls this method "very slow" (:64-66): it recomputes the
    Hamiltonian of the whole lattice per step; MetropolisHastingsSweep is the production path."""

    def __init__(self, number_of_update, spread):
        self._n, self._spread = int(number_of_update), float(spread)

    @classmethod
    def new(cls, number_of_update, spread):
        """None for invalid parameters (metropolis_hastings.rs:73-85)."""
        if number_of_update == 0 or spread <= 0.0 or spread >= 1.0:
            return None
        return cls(number_of_update, spread)

    def number_of_update(self):
        return self._n

    def spread(self):
        return self._spread

    def potential_next_element(self, state, rng):
        new = state.clone()
        seed, counter = _draw(rng)
        try:
            new._touch().metropolis_hits(seed, counter, self._spread, self._n, force_accept=True)
        except LqError as e:
            raise _wrap(e)
        return new


class MetropolisHastingsDiagnostic(MetropolisHastings):
    """metropolis_hastings.rs:160-290: the same with prob_replace_last / has_replace_last."""

    def __init__(self, number_of_update, spread):
        super().__init__(number_of_update, spread)
        self._prob_replace_last, self._has_replace_last = 0.0, False

    def prob_replace_last(self):
        return self._prob_replace_last

    def has_replace_last(self):
        return self._has_replace_last

    def next_element_default(self, state, rng):
        potential_next = self.potential_next_element(state, rng)
        proba = max(min(self.probability_of_replacement(state, potential_next), 1.0), 0.0)
        self._prob_replace_last = proba
        self._has_replace_last = _bernoulli(rng, proba)
        return potential_next if self._has_replace_last else state


class MetropolisHastingsDeltaDiagnostic(MonteCarlo):
    """metropolis_hastings.rs:300-417 (the README's method): every call proposes a change of ONE uniformly random link
    and accepts it on the local action difference delta_s_old_new_cmp (monte_carlo/mod.rs:324-334).

    `hits_per_call` = 1 restates the reference call for call.  Larger values run that many independent hits per call in
    one pair of launches (all on links of one random (direction, colour) class, so they do not see each other; hits
    that collide on a link are dropped): the Markov chain is the same random-scan Metropolis, 10^4-10^5 times faster
    per hit.  prob_replace_last() / has_replace_last() describe the last call: the mean acceptance probability of its
    hits and whether any was accepted."""

    def __init__(self, spread, rng, hits_per_call=1):
        self._spread, self._rng, self._hits = float(spread), rng, int(hits_per_call)
        self._prob_replace_last, self._has_replace_last = 0.0, False
        self.hits_performed_last, self.hits_accepted_last = 0, 0

    @classmethod
    def new(cls, spread, rng, hits_per_call=1):
        """None for an invalid spread (metropolis_hastings.rs:343-353)."""
        if spread <= 0.0 or spread >= 1.0 or hits_per_call < 1:
            return None
        return cls(spread, rng, hits_per_call)

    def prob_replace_last(self):
        return self._prob_replace_last

    def has_replace_last(self):
        return self._has_replace_last

    def spread(self):
        return self._spread

    def rng(self):
        return self._rng

    def rng_mut(self):
        return self._rng

    def rng_owned(self):
        return self._rng

    def next_element(self, state):
        seed, counter = _draw(self._rng)
        try:
            n_perf, n_acc, sum_p = state._touch().metropolis_hits(seed, counter, self._spread, self._hits)
        except LqError as e:
            raise _wrap(e)
        self.hits_performed_last, self.hits_accepted_last = n_perf, n_acc
        self._prob_replace_last = sum_p / max(n_perf, 1)
        self._has_replace_last = n_acc > 0
        return state
