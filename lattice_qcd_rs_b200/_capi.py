"""ctypes binding of include/lqcd_b200.h (the C ABI a Rust/cgo/JNI shim would bind).

`load()` opens the CUDA library built in-tree (lattice_qcd_rs_b200/liblqcd_b200.so).  There is no CPU fallback:
if the library is missing or no CUDA device is visible, the package raises.  (tests/emu.py re-uses `bind()` on a
host-emulation build of the same kernel bodies for the CPU CI -- test infrastructure, never loaded from here.)
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblqcd_b200.so")

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p

LQ_OK = 0
ERRORS = {
    -1: "LQ_E_BADARG", -2: "LQ_E_SIZE", -3: "LQ_E_CUDA", -4: "LQ_E_COMM", -5: "LQ_E_ODD_EXTENT",
    -6: "LQ_E_GAUSS_DIVERGED", -7: "LQ_E_ZERO_STEPS", -8: "LQ_E_NOSNAPSHOT", -9: "LQ_E_NODEVICE",
}
SYNC_SYNC, LEAP_LEAP, SYNC_LEAP, LEAP_SYNC, SYMPLECTIC = range(5)
OR_ROTATION, OR_REVERSE, OR_SU2_SUBGROUPS = 0, 1, 2
FLAG_PAULI3_FIXED, FLAG_NO_KICK_MERGE, FLAG_GAUSS_FUSED, FLAG_GENERIC_KERNELS, FLAG_UNIFORM_DIRECTION = 1, 2, 4, 8, 16
FLAG_GAUSS_TWO_PASS = 32
FLAG_FOLD_HALO_SYNC = 512  # decomposed contexts: halo synchronisation folded into the projection kernel (| 2048: MD chain too); off by default
INTEGRATOR_SYMPLECTIC_EULER, INTEGRATOR_OMELYAN = 0, 1
OMELYAN_LAMBDA = 0.1931833275037836  # second-order minimum-norm coefficient (Omelyan, Mryglod, Folk 2003)

HALO_FN = C.CFUNCTYPE(C.c_int, _vp, _vp, C.c_int)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, _vp, _dp, C.c_int)
ALLREDUCE_DEV_FN = C.CFUNCTYPE(C.c_int, _vp, _vp, C.c_int)


class LqComm(C.Structure):
    _fields_ = [("user", _vp), ("halo_exchange", HALO_FN), ("allreduce_sum", ALLREDUCE_FN),
                ("allreduce_sum_device", ALLREDUCE_DEV_FN)]


class LqError(RuntimeError):
    def __init__(self, code, where, detail=""):
        self.code = code
        self.name = ERRORS.get(code, str(code))
        super().__init__(f"{where}: {self.name} ({code}) {detail}".strip())


# every symbol include/lqcd_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "lq_strerror": (C.c_char_p, [C.c_int]),
    "lq_last_cuda_error": (C.c_char_p, []),
    "lq_version": (C.c_int, []),
    "lq_device_count": (C.c_int, [_ip]),
    "lq_ctx_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, _i64p, C.c_double, C.c_double, C.c_double]),
    "lq_ctx_create_dist": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, _i64p, _ip, _ip, C.c_double, C.c_double,
                                     C.c_double]),
    "lq_ctx_clone": (C.c_int, [_vp, C.POINTER(_vp)]),
    "lq_ctx_destroy": (C.c_int, [_vp]),
    "lq_set_flags": (C.c_int, [_vp, C.c_int]),
    "lq_get_flags": (C.c_int, [_vp, _ip]),
    "lq_set_beta": (C.c_int, [_vp, C.c_double]),
    "lq_sync": (C.c_int, [_vp]),
    "lq_stream": (C.c_int, [_vp, C.POINTER(_vp)]),
    "lq_num_sites": (C.c_int64, [_vp]),
    "lq_num_links": (C.c_int64, [_vp]),
    "lq_t": (C.c_int64, [_vp]),
    "lq_set_t": (C.c_int, [_vp, C.c_int64]),
    "lq_kernel_launches": (C.c_int64, [_vp]),
    "lq_links_upload": (C.c_int, [_vp, _dp, C.c_int64]),
    "lq_links_download": (C.c_int, [_vp, _dp, C.c_int64]),
    "lq_links_upload_begin": (C.c_int, [_vp, _dp, C.c_int64]),
    "lq_links_upload_commit": (C.c_int, [_vp]),
    "lq_links_download_begin": (C.c_int, [_vp, _dp, C.c_int64]),
    "lq_copies_wait": (C.c_int, [_vp]),
    "lq_efield_upload": (C.c_int, [_vp, _dp, C.c_int64]),
    "lq_efield_download": (C.c_int, [_vp, _dp, C.c_int64]),
    "lq_links_upload_device": (C.c_int, [_vp, _vp, C.c_int64]),
    "lq_links_download_device": (C.c_int, [_vp, _vp, C.c_int64]),
    "lq_efield_upload_device": (C.c_int, [_vp, _vp, C.c_int64]),
    "lq_efield_download_device": (C.c_int, [_vp, _vp, C.c_int64]),
    "lq_links_set_cold": (C.c_int, [_vp]),
    "lq_efield_set_zero": (C.c_int, [_vp]),
    "lq_links_set_random": (C.c_int, [_vp, C.c_uint64, C.c_uint64]),
    "lq_plaquette_sum": (C.c_int, [_vp, _dp]),
    "lq_average_trace_plaquette": (C.c_int, [_vp, _dp]),
    "lq_hamiltonian_links": (C.c_int, [_vp, _dp]),
    "lq_hamiltonian_efield": (C.c_int, [_vp, _dp]),
    "lq_hamiltonian_total": (C.c_int, [_vp, _dp]),
    "lq_clover": (C.c_int, [_vp, C.c_int, C.c_int, _dp, C.c_int64]),
    "lq_f_mu_nu": (C.c_int, [_vp, C.c_int, C.c_int, _dp, C.c_int64]),
    "lq_magnetic_field": (C.c_int, [_vp, C.c_int, _dp, C.c_int64]),
    "lq_staples": (C.c_int, [_vp, _dp, C.c_int64]),
    "lq_force": (C.c_int, [_vp, _dp, C.c_int64]),
    "lq_efield_step": (C.c_int, [_vp, C.c_double]),
    "lq_link_step": (C.c_int, [_vp, C.c_double, C.c_int]),
    "lq_integrate": (C.c_int, [_vp, C.c_int, C.c_double]),
    "lq_symplectic_n": (C.c_int, [_vp, C.c_double, C.c_int64]),
    "lq_leapfrog_n": (C.c_int, [_vp, C.c_double, C.c_int64]),
    "lq_set_integrator": (C.c_int, [_vp, C.c_int, C.c_double, C.c_int]),
    "lq_md_n": (C.c_int, [_vp, C.c_double, C.c_int64]),
    "lq_reunitarize": (C.c_int, [_vp]),
    "lq_momenta_refresh": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_double]),
    "lq_gauss_field": (C.c_int, [_vp, _dp, C.c_int64]),
    "lq_gauss_sum_div": (C.c_int, [_vp, _dp]),
    "lq_gauss_project_step": (C.c_int, [_vp]),
    "lq_gauss_project": (C.c_int, [_vp, C.c_int64, _i64p]),
    "lq_sweep_heatbath": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_double]),
    "lq_sweep_overrelax": (C.c_int, [_vp, C.c_int]),
    "lq_sweep_metropolis": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_double, C.c_int, _i64p, _dp]),
    "lq_metropolis_hits": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_double, C.c_int64, C.c_int, _i64p, _i64p, _dp]),
    "lq_snapshot": (C.c_int, [_vp]),
    "lq_restore": (C.c_int, [_vp]),
    "lq_hmc_trajectory": (C.c_int, [_vp, C.c_double, C.c_int64, C.c_uint64, C.c_uint64, C.c_double, C.c_int, C.c_int,
                                    _dp, _dp, _dp, _ip, _i64p]),
    "lq_set_comm": (C.c_int, [_vp, C.POINTER(LqComm)]),
    "lq_set_stream": (C.c_int, [_vp, _vp]),
    "lq_is_decomposed": (C.c_int, [_vp, C.c_int]),
    "lq_halo_bytes": (C.c_int, [_vp, C.c_int, C.c_int, _i64p]),
    "lq_halo_pack": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int64]),
    "lq_halo_unpack": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int64]),
    "lq_halo_invalidate": (C.c_int, [_vp, C.c_int]),
    "lq_p2p_export": (C.c_int, [_vp, _vp, C.c_int64]),
    "lq_p2p_attach": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _ip, _ip]),
    "lq_p2p_enabled": (C.c_int, [_vp]),
    "lq_p2p_exchanges": (C.c_int64, [_vp]),
    "lq_profile_enable": (C.c_int, [_vp, C.c_int]),
    "lq_profile_reset": (C.c_int, [_vp]),
    "lq_profile_get": (C.c_int, [_vp, C.c_int, _i64p, _dp]),
    "lq_measure_peaks": (C.c_int, [_vp, _dp, _dp]),
}
PROF = {"efield_link_step": 0, "efield_step": 1, "link_step": 2, "plaquette": 3, "gauss_field": 4, "gauss_step": 5,
        "heatbath": 6, "overrelax": 7, "metropolis": 8, "reunitarize": 9, "momenta": 10, "efield_energy": 11,
        "gauss_div": 12, "copy": 13}


def bind(path):
    """dlopen `path` and declare every symbol of the C ABI (raises AttributeError on a missing export)."""
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_LIB = None


def load():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m lattice_qcd_rs_b200.build` (nvcc, sm_100a). "
                "There is no CPU fallback.")
        _LIB = bind(LIB_PATH)
    return _LIB


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _p(a):
    return a.ctypes.data_as(_dp)


class Context:
    """Owns one lq_ctx.  Thin, 1:1 with the C ABI; arrays are numpy f64 in the reference AoS layouts."""

    def __init__(self, D, extent, a=1.0, beta=1.0, CA=3.0, device=0, lib=None, proc_grid=None, rank_coord=None):
        self.lib = lib if lib is not None else load()
        if np.isscalar(extent):
            extent = [int(extent)] * D
        if len(extent) != D:
            raise LqError(-1, "Context", "len(extent) != D")
        self.D = int(D)
        self.global_extent = [int(e) for e in extent]
        self.a, self.beta, self.CA = float(a), float(beta), float(CA)
        ext = (C.c_int64 * D)(*self.global_extent)
        h = _vp()
        if proc_grid is None:
            rc = self.lib.lq_ctx_create(C.byref(h), device, D, ext, self.a, self.beta, self.CA)
            self.proc_grid, self.rank_coord = [1] * D, [0] * D
        else:
            pg = (C.c_int * D)(*proc_grid)
            rcoord = (C.c_int * D)(*rank_coord)
            rc = self.lib.lq_ctx_create_dist(C.byref(h), device, D, ext, pg, rcoord, self.a, self.beta, self.CA)
            self.proc_grid, self.rank_coord = list(proc_grid), list(rank_coord)
        self._h = h
        self._check(rc, "lq_ctx_create")
        self.extent = [e // p for e, p in zip(self.global_extent, self.proc_grid)]
        self.ns = int(self.lib.lq_num_sites(self._h))
        self.nl = int(self.lib.lq_num_links(self._h))
        self._comm_keepalive = None
        self._borrowed = False

    # -- plumbing
    def _check(self, rc, where):
        if rc != LQ_OK:
            detail = ""
            if rc == -3:
                detail = self.lib.lq_last_cuda_error().decode()
            raise LqError(rc, where, detail)

    def clone(self):
        """Device-to-device copy of this context (lattice, beta, flags, t, links, E-field)."""
        h = _vp()
        self._check(self.lib.lq_ctx_clone(self._h, C.byref(h)), "lq_ctx_clone")
        new = object.__new__(Context)
        new.__dict__.update({k: v for k, v in self.__dict__.items() if k != "_h"})
        new._h = h
        new._borrowed = False
        return new

    def borrowed(self, handle):
        """A non-owning view of another lq_ctx* with this context's geometry (the handle a halo callback receives: this
        context itself or a clone of it); closing the view does not destroy the context."""
        if handle is None or handle == self._h.value:
            return self
        v = object.__new__(Context)
        v.__dict__.update({k: x for k, x in self.__dict__.items() if k != "_h"})
        v._h = _vp(handle)
        v._borrowed = True
        return v

    def close(self):
        if getattr(self, "_h", None):
            if not getattr(self, "_borrowed", False):
                self.lib.lq_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def sync(self):
        self._check(self.lib.lq_sync(self._h), "lq_sync")

    def stream(self):
        s = _vp()
        self._check(self.lib.lq_stream(self._h, C.byref(s)), "lq_stream")
        return s.value or 0

    def set_stream(self, cuda_stream):
        self._check(self.lib.lq_set_stream(self._h, _vp(cuda_stream)), "lq_set_stream")

    def set_flags(self, flags):
        self._check(self.lib.lq_set_flags(self._h, int(flags)), "lq_set_flags")

    def set_beta(self, beta):
        self.beta = float(beta)
        self._check(self.lib.lq_set_beta(self._h, self.beta), "lq_set_beta")

    @property
    def t(self):
        return int(self.lib.lq_t(self._h))

    def set_t(self, t):
        self._check(self.lib.lq_set_t(self._h, int(t)), "lq_set_t")

    @property
    def kernel_launches(self):
        return int(self.lib.lq_kernel_launches(self._h))

    def set_comm(self, halo_exchange, allreduce_sum, allreduce_sum_device=None):
        """halo_exchange(ctx_handle, which) -> int (the handle is the lq_ctx* the library asks the refresh for: this
        context or a clone of it), allreduce_sum(numpy view of n doubles) -> int,
        allreduce_sum_device(pointer, n) -> int (optional: in-place sum in the library's result buffer)."""
        def _halo(user, ctx, which):
            try:
                return int(halo_exchange(ctx, which) or 0)
            except Exception:  # never unwind across the ABI
                import traceback
                traceback.print_exc()
                return -4

        def _allr(user, vals, n):
            try:
                arr = np.ctypeslib.as_array(vals, shape=(n,))
                return int(allreduce_sum(arr) or 0)
            except Exception:
                import traceback
                traceback.print_exc()
                return -4

        def _allr_dev(user, ptr, n):
            try:
                return int(allreduce_sum_device(ptr, n) or 0)
            except Exception:
                import traceback
                traceback.print_exc()
                return -4

        comm = LqComm(None, HALO_FN(_halo), ALLREDUCE_FN(_allr),
                      ALLREDUCE_DEV_FN(_allr_dev) if allreduce_sum_device is not None else ALLREDUCE_DEV_FN())
        self._comm_keepalive = comm
        self._check(self.lib.lq_set_comm(self._h, C.byref(comm)), "lq_set_comm")

    # -- marshalling
    def links_upload(self, U):
        U = _f64(U)
        self._check(self.lib.lq_links_upload(self._h, _p(U), U.size // 18), "lq_links_upload")

    def links_download(self, out=None):
        U = np.empty((self.nl, 18)) if out is None else out
        assert U.dtype == np.float64 and U.flags["C_CONTIGUOUS"] and U.size == self.nl * 18
        self._check(self.lib.lq_links_download(self._h, _p(U), self.nl), "lq_links_download")
        return U

    # pipelined marshalling: the arrays must stay alive and untouched until copies_wait() (use pinned memory for copies
    # that really run behind the kernels)
    def links_upload_begin(self, U):
        assert U.dtype == np.float64 and U.flags["C_CONTIGUOUS"] and U.size == self.nl * 18
        self._inflight_up = U
        self._check(self.lib.lq_links_upload_begin(self._h, _p(U), self.nl), "lq_links_upload_begin")

    def links_upload_commit(self):
        self._check(self.lib.lq_links_upload_commit(self._h), "lq_links_upload_commit")

    def links_download_begin(self, out):
        assert out.dtype == np.float64 and out.flags["C_CONTIGUOUS"] and out.size == self.nl * 18
        self._inflight_down = out
        self._check(self.lib.lq_links_download_begin(self._h, _p(out), self.nl), "lq_links_download_begin")

    def copies_wait(self):
        self._check(self.lib.lq_copies_wait(self._h), "lq_copies_wait")
        self._inflight_up = self._inflight_down = None

    def efield_upload(self, E):
        E = _f64(E)
        self._check(self.lib.lq_efield_upload(self._h, _p(E), E.size // 8), "lq_efield_upload")

    def efield_download(self, out=None):
        E = np.empty((self.nl, 8)) if out is None else out
        assert E.dtype == np.float64 and E.flags["C_CONTIGUOUS"] and E.size == self.nl * 8
        self._check(self.lib.lq_efield_download(self._h, _p(E), self.nl), "lq_efield_download")
        return E

    def links_upload_device(self, ptr, n_links):
        self._check(self.lib.lq_links_upload_device(self._h, _vp(ptr), n_links), "lq_links_upload_device")

    def links_download_device(self, ptr, n_links):
        self._check(self.lib.lq_links_download_device(self._h, _vp(ptr), n_links), "lq_links_download_device")

    def efield_upload_device(self, ptr, n_links):
        self._check(self.lib.lq_efield_upload_device(self._h, _vp(ptr), n_links), "lq_efield_upload_device")

    def efield_download_device(self, ptr, n_links):
        self._check(self.lib.lq_efield_download_device(self._h, _vp(ptr), n_links), "lq_efield_download_device")

    def links_set_cold(self):
        self._check(self.lib.lq_links_set_cold(self._h), "lq_links_set_cold")

    def efield_set_zero(self):
        self._check(self.lib.lq_efield_set_zero(self._h), "lq_efield_set_zero")

    def links_set_random(self, seed, counter=0):
        self._check(self.lib.lq_links_set_random(self._h, seed, counter), "lq_links_set_random")

    # -- observables
    def plaquette_sum(self):
        out = np.empty(2)
        self._check(self.lib.lq_plaquette_sum(self._h, _p(out)), "lq_plaquette_sum")
        return complex(out[0], out[1])

    def average_trace_plaquette(self):
        out = np.empty(2)
        self._check(self.lib.lq_average_trace_plaquette(self._h, _p(out)), "lq_average_trace_plaquette")
        return complex(out[0], out[1])

    def _scalar(self, fn, name):
        v = C.c_double(0)
        self._check(fn(self._h, C.byref(v)), name)
        return v.value

    def hamiltonian_links(self):
        return self._scalar(self.lib.lq_hamiltonian_links, "lq_hamiltonian_links")

    def hamiltonian_efield(self):
        return self._scalar(self.lib.lq_hamiltonian_efield, "lq_hamiltonian_efield")

    def hamiltonian_total(self):
        return self._scalar(self.lib.lq_hamiltonian_total, "lq_hamiltonian_total")

    # -- field-strength observables (signed directions: +(d+1) / -(d+1))
    def clover(self, sdir_i, sdir_j):
        out = np.empty((self.ns, 18))
        self._check(self.lib.lq_clover(self._h, sdir_i, sdir_j, _p(out), self.ns), "lq_clover")
        return out

    def f_mu_nu(self, dir_i, dir_j):
        out = np.empty((self.ns, 18))
        self._check(self.lib.lq_f_mu_nu(self._h, dir_i, dir_j, _p(out), self.ns), "lq_f_mu_nu")
        return out

    def magnetic_field(self, direction):
        out = np.empty((self.ns, 18))
        self._check(self.lib.lq_magnetic_field(self._h, direction, _p(out), self.ns), "lq_magnetic_field")
        return out

    # -- molecular dynamics
    def staples(self):
        out = np.empty((self.nl, 18))
        self._check(self.lib.lq_staples(self._h, _p(out), self.nl), "lq_staples")
        return out

    def force(self):
        out = np.empty((self.nl, 8))
        self._check(self.lib.lq_force(self._h, _p(out), self.nl), "lq_force")
        return out

    def efield_step(self, dt):
        self._check(self.lib.lq_efield_step(self._h, dt), "lq_efield_step")

    def link_step(self, dt, use_exp=False):
        self._check(self.lib.lq_link_step(self._h, dt, int(use_exp)), "lq_link_step")

    def integrate(self, kind, dt):
        self._check(self.lib.lq_integrate(self._h, int(kind), dt), "lq_integrate")

    def symplectic_n(self, dt, n):
        self._check(self.lib.lq_symplectic_n(self._h, dt, int(n)), "lq_symplectic_n")

    def leapfrog_n(self, dt, n):
        self._check(self.lib.lq_leapfrog_n(self._h, dt, int(n)), "lq_leapfrog_n")

    def set_integrator(self, kind=INTEGRATOR_SYMPLECTIC_EULER, lam=OMELYAN_LAMBDA, use_exp=False):
        """What md_n / hmc_trajectory integrate with; the default (0, -, False) is the reference's symplectic Euler."""
        self._check(self.lib.lq_set_integrator(self._h, int(kind), float(lam), int(use_exp)), "lq_set_integrator")

    def md_n(self, dt, n):
        self._check(self.lib.lq_md_n(self._h, dt, int(n)), "lq_md_n")

    def reunitarize(self):
        self._check(self.lib.lq_reunitarize(self._h), "lq_reunitarize")

    # -- momenta + Gauss
    def momenta_refresh(self, seed, counter, sigma=None):
        sigma = 0.5 / self.beta if sigma is None else sigma
        self._check(self.lib.lq_momenta_refresh(self._h, seed, counter, sigma), "lq_momenta_refresh")

    def gauss_field(self):
        out = np.empty((self.ns, 18))
        self._check(self.lib.lq_gauss_field(self._h, _p(out), self.ns), "lq_gauss_field")
        return out

    def gauss_sum_div(self):
        return self._scalar(self.lib.lq_gauss_sum_div, "lq_gauss_sum_div")

    def gauss_project_step(self):
        self._check(self.lib.lq_gauss_project_step(self._h), "lq_gauss_project_step")

    def gauss_project(self, max_steps=0):
        it = C.c_int64(0)
        self._check(self.lib.lq_gauss_project(self._h, max_steps, C.byref(it)), "lq_gauss_project")
        return it.value

    # -- sweeps
    def sweep_heatbath(self, seed, counter, coupling_scale=1.0):
        self._check(self.lib.lq_sweep_heatbath(self._h, seed, counter, coupling_scale), "lq_sweep_heatbath")

    def sweep_overrelax(self, kind):
        self._check(self.lib.lq_sweep_overrelax(self._h, int(kind)), "lq_sweep_overrelax")

    def sweep_metropolis(self, seed, counter, spread=0.1, n_update=1):
        na, sp = C.c_int64(0), C.c_double(0)
        self._check(self.lib.lq_sweep_metropolis(self._h, seed, counter, spread, n_update, C.byref(na), C.byref(sp)),
                    "lq_sweep_metropolis")
        return na.value, sp.value

    def metropolis_hits(self, seed, counter, spread=0.1, n_hits=1, force_accept=False):
        """n_hits random single-link Metropolis hits (MetropolisHastingsDeltaDiagnostic, one hit = the reference's call).
        Returns (n_performed, n_accepted, sum of acceptance probabilities)."""
        npf, na, sp = C.c_int64(0), C.c_int64(0), C.c_double(0)
        self._check(self.lib.lq_metropolis_hits(self._h, seed, counter, spread, int(n_hits), int(force_accept),
                                                C.byref(npf), C.byref(na), C.byref(sp)), "lq_metropolis_hits")
        return npf.value, na.value, sp.value

    # -- HMC
    def snapshot(self):
        self._check(self.lib.lq_snapshot(self._h), "lq_snapshot")

    def restore(self):
        self._check(self.lib.lq_restore(self._h), "lq_restore")

    def hmc_trajectory(self, dt, n_steps, seed, counter, sigma=None, use_current_e=False, do_project=True):
        sigma = 0.5 / self.beta if sigma is None else sigma
        h0, h1, p = C.c_double(0), C.c_double(0), C.c_double(0)
        acc, gs = C.c_int(0), C.c_int64(0)
        self._check(self.lib.lq_hmc_trajectory(self._h, dt, int(n_steps), seed, counter, sigma, int(use_current_e),
                                               int(do_project), C.byref(h0), C.byref(h1), C.byref(p), C.byref(acc),
                                               C.byref(gs)), "lq_hmc_trajectory")
        return dict(h_old=h0.value, h_new=h1.value, prob=p.value, accepted=bool(acc.value), gauss_steps=gs.value)

    # -- measurement
    def profile_enable(self, on=True):
        self._check(self.lib.lq_profile_enable(self._h, int(on)), "lq_profile_enable")

    def profile_reset(self):
        self._check(self.lib.lq_profile_reset(self._h), "lq_profile_reset")

    def profile_get(self, kernel):
        n, ms = C.c_int64(0), C.c_double(0)
        self._check(self.lib.lq_profile_get(self._h, PROF[kernel], C.byref(n), C.byref(ms)), "lq_profile_get")
        return n.value, ms.value

    def measure_peaks(self):
        """(f64 FMA TFLOP/s, streaming-copy GB/s) measured on this context's device now."""
        f, b = C.c_double(0), C.c_double(0)
        self._check(self.lib.lq_measure_peaks(self._h, C.byref(f), C.byref(b)), "lq_measure_peaks")
        return f.value, b.value

    # -- halos
    def is_decomposed(self, d):
        return bool(self.lib.lq_is_decomposed(self._h, d))

    def halo_bytes(self, which, d):
        b = C.c_int64(0)
        self._check(self.lib.lq_halo_bytes(self._h, which, d, C.byref(b)), "lq_halo_bytes")
        return b.value

    def halo_pack(self, which, d, side, ptr, nbytes):
        self._check(self.lib.lq_halo_pack(self._h, which, d, side, _vp(ptr), nbytes), "lq_halo_pack")

    def halo_unpack(self, which, d, side, ptr, nbytes):
        self._check(self.lib.lq_halo_unpack(self._h, which, d, side, _vp(ptr), nbytes), "lq_halo_unpack")

    # -- peer-to-peer transport
    def p2p_export(self):
        buf = C.create_string_buffer(9 * 64)
        self._check(self.lib.lq_p2p_export(self._h, buf, 9 * 64), "lq_p2p_export")
        return buf.raw

    def p2p_attach(self, peer_handles, offsets, peer_index):
        """peer_handles: list of 576-byte blobs (one per unique peer); offsets: list of D-vectors; peer_index: list."""
        blob = b"".join(peer_handles)
        n_nb = len(offsets)
        off = (C.c_int * (n_nb * self.D))(*[int(v) for o in offsets for v in o])
        pi = (C.c_int * n_nb)(*[int(v) for v in peer_index])
        self._check(self.lib.lq_p2p_attach(self._h, len(peer_handles), C.c_char_p(blob), n_nb, off, pi), "lq_p2p_attach")

    @property
    def p2p_enabled(self):
        return bool(self.lib.lq_p2p_enabled(self._h))

    @property
    def p2p_exchanges(self):
        return int(self.lib.lq_p2p_exchanges(self._h))

    def halo_invalidate(self, which):
        self._check(self.lib.lq_halo_invalidate(self._h, which), "lq_halo_invalidate")
