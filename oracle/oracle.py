"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (lattice_qcd_rs_b200) never does.

Arrays are numpy float64 in the reference's AoS layouts:
  links  (Nl, 18): link index = site*D + dir, 3x3 complex column-major (re, im)   field.rs:584-586
  efield (Nl, 8)                                                                  field.rs:1025-1027
Helper `to_c(U)` / `from_c(M)` convert to/from complex (Nl, 3, 3) row/col indexed arrays.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_u32p = C.POINTER(C.c_uint32)


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def to_c(U):
    """(N,18) f64 AoS -> (N,3,3) complex, [n, row, col]."""
    U = np.asarray(U, dtype=np.float64).reshape(-1, 3, 3, 2)  # [n, col, row, reim]
    return (U[..., 0] + 1j * U[..., 1]).transpose(0, 2, 1)


def from_c(M):
    """(N,3,3) complex [n,row,col] -> (N,18) f64 AoS."""
    M = np.asarray(M, dtype=np.complex128).reshape(-1, 3, 3).transpose(0, 2, 1)  # [n, col, row]
    out = np.empty(M.shape + (2,), dtype=np.float64)
    out[..., 0] = M.real
    out[..., 1] = M.imag
    return np.ascontiguousarray(out.reshape(-1, 18))


class Oracle:
    """One lattice geometry (D, extents, spacing a) + coupling (beta, CA)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            path = _build.build()
            cls._lib = C.CDLL(path)
            L = cls._lib
            L.lqo_hamiltonian_links.restype = C.c_double
            L.lqo_hamiltonian_efield.restype = C.c_double
            L.lqo_gauss_sum_div.restype = C.c_double
            L.lqo_delta_s.restype = C.c_double
            L.lqo_project_to_gauss.restype = C.c_int64
            L.lqo_hmc_trajectory.restype = C.c_int64
            L.lqo_num_threads.restype = C.c_int
        return cls._lib

    def __init__(self, D, ext, a=1.0, beta=1.0, CA=3.0):
        if np.isscalar(ext):
            ext = [int(ext)] * D
        assert len(ext) == D
        self.D = int(D)
        self.ext = [int(e) for e in ext]
        self._ext = (C.c_int64 * D)(*self.ext)
        self.a = float(a)
        self.beta = float(beta)
        self.CA = float(CA)
        self.ns = int(np.prod(self.ext))
        self.nl = self.ns * self.D
        self.L = self.lib()

    # -- helpers
    def _g(self):
        return (C.c_int(self.D), self._ext, C.c_double(self.a))

    def site_index(self, x):
        s, stride = 0, 1
        for k in range(self.D):
            s += (x[k] % self.ext[k]) * stride
            stride *= self.ext[k]
        return s

    @staticmethod
    def sdir(i, positive=True):
        return (i + 1) if positive else -(i + 1)

    def cold_links(self):
        U = np.zeros((self.nl, 18))
        U[:, 0] = U[:, 8] = U[:, 16] = 1.0
        return U

    def cold_efield(self):
        return np.zeros((self.nl, 8))

    def num_threads(self):
        return self.L.lqo_num_threads()

    FLAG_PAULI3_FIXED = 1
    FLAG_UNIFORM_DIRECTION = 16

    def set_flags(self, flags):
        """Process-global behaviour flags (0 = reference as coded; 1 = true sigma_3; 16 = sphere-uniform heat-bath
        direction)."""
        self.L.lqo_set_flags(C.c_int(flags))

    def set_num_threads(self, n):
        self.L.lqo_set_num_threads(C.c_int(n))

    # -- algebra (static-like)
    def generator(self, a):
        out = np.empty(18)
        self.L.lqo_generator(C.c_int(a), _p(out))
        return to_c(out)[0]

    def adjoint_to_matrix(self, e8):
        e8 = np.ascontiguousarray(e8, dtype=np.float64)
        out = np.empty(18)
        self.L.lqo_adjoint_to_matrix(_p(e8), _p(out))
        return to_c(out)[0]

    def su3_exp_i(self, e8):
        e8 = np.ascontiguousarray(e8, dtype=np.float64)
        out = np.empty(18)
        self.L.lqo_su3_exp_i(_p(e8), _p(out))
        return to_c(out)[0]

    def orthonormalize(self, M):
        a = from_c(M)[0].copy()
        out = np.empty(18)
        self.L.lqo_orthonormalize(_p(a), _p(out))
        return to_c(out)[0]

    def svd3(self, M):
        a = from_c(M)[0].copy()
        u, v, s = np.empty(18), np.empty(18), np.empty(3)
        self.L.lqo_svd3(_p(a), _p(u), _p(s), _p(v))
        return to_c(u)[0], s, to_c(v)[0]

    def overrelax_link(self, Ulink, staple, kind):
        u, a, out = from_c(Ulink)[0].copy(), from_c(staple)[0].copy(), np.empty(18)
        self.L.lqo_overrelax_link(_p(u), _p(a), C.c_int(kind), _p(out))
        return to_c(out)[0]

    def delta_s(self, staple, new, old):
        return self.L.lqo_delta_s(_p(from_c(staple)[0].copy()), _p(from_c(new)[0].copy()), _p(from_c(old)[0].copy()),
                                  C.c_double(self.beta), C.c_double(self.CA))

    # -- rng
    def philox_block(self, ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        self.L.lqo_philox_block(c, k, o)
        return list(o)

    def stream_uniform01(self, seed, counter, idx, n):
        out = np.empty(n)
        self.L.lqo_stream_uniform01(C.c_uint64(seed), C.c_uint64(counter), C.c_uint64(idx), C.c_int64(n), _p(out))
        return out

    def random_su3_close_to_unity(self, seed, counter, idx, spread):
        out = np.empty(18)
        self.L.lqo_random_su3_close_to_unity(C.c_uint64(seed), C.c_uint64(counter), C.c_uint64(idx),
                                             C.c_double(spread), _p(out))
        return to_c(out)[0]

    def heat_bath_norm_samples(self, seed, counter, param, n):
        out = np.empty(n)
        self.L.lqo_heat_bath_norm_samples(C.c_uint64(seed), C.c_uint64(counter), C.c_double(param), C.c_int64(n),
                                          _p(out))
        return out

    # -- local quantities
    def _local(self, fn, U, site, si, sj):
        out = np.empty(18)
        fn(*self._g(), _p(U), C.c_int64(site), C.c_int(si), C.c_int(sj), _p(out))
        return to_c(out)[0]

    def sij(self, U, site, si, sj):
        return self._local(self.L.lqo_sij, U, site, si, sj)

    def pij(self, U, site, si, sj):
        return self._local(self.L.lqo_pij, U, site, si, sj)

    def clover(self, U, site, si, sj):
        return self._local(self.L.lqo_clover, U, site, si, sj)

    def f_mu_nu(self, U, site, si, sj):
        return self._local(self.L.lqo_f_mu_nu, U, site, si, sj)

    def magnetic_field(self, U, site, d):
        out = np.empty(18)
        self.L.lqo_magnetic_field(*self._g(), _p(U), C.c_int64(site), C.c_int(d), _p(out))
        return to_c(out)[0]

    def staples(self, U):
        out = np.empty((self.nl, 18))
        self.L.lqo_staples(*self._g(), _p(U), _p(out))
        return out

    # -- observables
    def plaquette_sum(self, U):
        out = np.empty(2)
        self.L.lqo_plaquette_sum(*self._g(), _p(U), _p(out))
        return complex(out[0], out[1])

    def average_trace_plaquette(self, U):
        """field.rs:775-804: sum / (Ns * D(D-1)/2)."""
        return self.plaquette_sum(U) / (self.ns * (self.D * (self.D - 1)) // 2)

    def hamiltonian_links(self, U):
        return self.L.lqo_hamiltonian_links(*self._g(), _p(U), C.c_double(self.beta), C.c_double(self.CA))

    def hamiltonian_efield(self, E):
        return self.L.lqo_hamiltonian_efield(*self._g(), _p(E), C.c_double(self.beta))

    def hamiltonian_total(self, U, E):
        return self.hamiltonian_links(U) + self.hamiltonian_efield(E)

    # -- molecular dynamics (return new arrays; inputs untouched)
    def force(self, U, literal=True):
        F = np.empty((self.nl, 8))
        self.L.lqo_force(*self._g(), _p(U), C.c_double(self.CA), _p(F), C.c_int(int(literal)))
        return F

    def efield_step(self, U, E, dt, literal=True):
        E = E.copy()
        self.L.lqo_efield_step(*self._g(), _p(U), _p(E), C.c_double(dt), C.c_double(self.CA), C.c_int(int(literal)))
        return E

    def link_step(self, U, E, dt):
        U = U.copy()
        self.L.lqo_link_step(*self._g(), _p(U), _p(E), C.c_double(dt), C.c_double(self.CA))
        return U

    def link_step_exp(self, U, E, dt):
        U = U.copy()
        self.L.lqo_link_step_exp(*self._g(), _p(U), _p(E), C.c_double(dt), C.c_double(self.CA))
        return U

    KINDS = {"sync_sync": 0, "leap_leap": 1, "sync_leap": 2, "leap_sync": 3, "symplectic": 4}

    def integrate(self, U, E, kind, dt, n=1, literal=True):
        U, E = U.copy(), E.copy()
        k = self.KINDS[kind] if isinstance(kind, str) else int(kind)
        self.L.lqo_integrate(*self._g(), _p(U), _p(E), C.c_int(k), C.c_double(dt), C.c_int64(n), C.c_double(self.CA),
                             C.c_int(int(literal)))
        return U, E

    def md_n(self, U, E, dt, n, kind=0, lam=0.1931833275037836, use_exp=False, literal=True):
        """Integrator options beyond the crate's, composed from the two reference updates
        integrate_efield (integrator/mod.rs:240-254) and integrate_link (:216-233) / its exponential form:
        kind 0: E(dt/2) U(dt) E(dt/2) per step; kind 1 (Omelyan): E(l dt) U(dt/2) E((1-2l) dt) U(dt/2) E(l dt);
        adjacent E kicks of consecutive steps merged (one kick of the summed length, as the library does)."""
        ustep = self.link_step_exp if use_exp else self.link_step
        if kind == 0:
            E = self.efield_step(U, E, dt / 2.0, literal)
            for k in range(n):
                U = ustep(U, E, dt)
                E = self.efield_step(U, E, dt if k + 1 < n else dt / 2.0, literal)
        else:
            E = self.efield_step(U, E, lam * dt, literal)
            for k in range(n):
                U = ustep(U, E, dt / 2.0)
                E = self.efield_step(U, E, (1.0 - 2.0 * lam) * dt, literal)
                U = ustep(U, E, dt / 2.0)
                E = self.efield_step(U, E, 2.0 * lam * dt if k + 1 < n else lam * dt, literal)
        return U, E

    def leapfrog_n(self, U, E, dt, n, literal=True):
        U, E = U.copy(), E.copy()
        self.L.lqo_leapfrog_n(*self._g(), _p(U), _p(E), C.c_double(dt), C.c_int64(n), C.c_double(self.CA),
                              C.c_int(int(literal)))
        return U, E

    # -- Gauss
    def gauss_field(self, U, E):
        out = np.empty((self.ns, 18))
        self.L.lqo_gauss_field(*self._g(), _p(U), _p(E), _p(out))
        return out

    def gauss_sum_div(self, U, E):
        return self.L.lqo_gauss_sum_div(*self._g(), _p(U), _p(E))

    def project_to_gauss_step(self, U, E):
        E = E.copy()
        self.L.lqo_project_to_gauss_step(*self._g(), _p(U), _p(E))
        return E

    def project_to_gauss(self, U, E, max_steps=1 << 20):
        E = E.copy()
        it = self.L.lqo_project_to_gauss(*self._g(), _p(U), _p(E), C.c_int64(max_steps))
        return E, it

    # -- reprojection
    def normalize_links(self, U):
        U = U.copy()
        self.L.lqo_normalize_links(*self._g(), _p(U))
        return U

    # -- start configs
    def links_random(self, seed, counter=0):
        U = np.empty((self.nl, 18))
        self.L.lqo_links_random(*self._g(), _p(U), C.c_uint64(seed), C.c_uint64(counter))
        return U

    def momenta_refresh(self, seed, counter, sigma=None):
        E = np.empty((self.nl, 8))
        sigma = 0.5 / self.beta if sigma is None else sigma
        self.L.lqo_momenta_refresh(*self._g(), _p(E), C.c_uint64(seed), C.c_uint64(counter), C.c_double(sigma))
        return E

    # -- sweeps (order: 0 sequential = reference, 1 checkerboard = CUDA order)
    def sweep_heatbath(self, U, seed, counter, order=1, per_link=True, coupling_scale=1.0):
        U = U.copy()
        self.L.lqo_sweep_heatbath(*self._g(), _p(U), C.c_double(self.beta), C.c_double(coupling_scale), C.c_int(order),
                                  C.c_uint64(seed), C.c_uint64(counter), C.c_int(int(per_link)))
        return U

    def sweep_overrelax(self, U, kind, order=1):
        U = U.copy()
        self.L.lqo_sweep_overrelax(*self._g(), _p(U), C.c_int(kind), C.c_int(order))
        return U

    def sweep_metropolis(self, U, seed, counter, n_update=1, spread=0.1, order=1, per_link=True):
        U = U.copy()
        na, sp = C.c_int64(0), C.c_double(0)
        self.L.lqo_sweep_metropolis(*self._g(), _p(U), C.c_double(self.beta), C.c_double(self.CA), C.c_int(n_update),
                                    C.c_double(spread), C.c_int(order), C.c_uint64(seed), C.c_uint64(counter),
                                    C.c_int(int(per_link)), C.byref(na), C.byref(sp))
        return U, na.value, sp.value

    def metropolis_single_link(self, U, seed, counter, spread, n_hits):
        U = U.copy()
        na, sp = C.c_int64(0), C.c_double(0)
        self.L.lqo_metropolis_single_link(*self._g(), _p(U), C.c_double(self.beta), C.c_double(self.CA),
                                          C.c_double(spread), C.c_int64(n_hits), C.c_uint64(seed), C.c_uint64(counter),
                                          C.byref(na), C.byref(sp))
        return U, na.value, sp.value

    # -- HMC
    def hmc_trajectory(self, U, dt, n_steps, seed, counter, E=None, do_project=True, sigma=None, literal=True):
        """Returns dict(U, E, h_old, h_new, prob, accepted, gauss_steps)."""
        U = U.copy()
        sigma = 0.5 / self.beta if sigma is None else sigma
        use_e = E is not None
        Eb = E.copy() if use_e else np.empty((self.nl, 8))
        h0, h1, p, acc = C.c_double(0), C.c_double(0), C.c_double(0), C.c_int(0)
        it = self.L.lqo_hmc_trajectory(*self._g(), _p(U), _p(Eb), C.c_int(int(use_e)), C.c_int(int(do_project)),
                                       C.c_double(self.beta), C.c_double(self.CA), C.c_double(sigma), C.c_double(dt),
                                       C.c_int64(n_steps), C.c_uint64(seed), C.c_uint64(counter), C.c_int(int(literal)),
                                       C.byref(h0), C.byref(h1), C.byref(p), C.byref(acc))
        return dict(U=U, E=Eb, h_old=h0.value, h_new=h1.value, prob=p.value, accepted=bool(acc.value), gauss_steps=it)
