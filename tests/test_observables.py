"""Field-strength observables adjacent to the update path: clover, f_mu_nu, magnetic_field
(field.rs:807-882) as kernels over every site, against the reference's own known answers (`magnetic_field` test,
field.rs:1580-1711) and against the oracle site by site."""
import numpy as np
import pytest

from oracle.oracle import Oracle, from_c, to_c
from tests.conftest import SEED_RNG

EPS = 1e-12


@pytest.fixture(params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "emu":
        from tests import emu
        return emu.context
    import torch
    assert torch.cuda.is_available(), "the gpu-marked tests need a CUDA device (no CPU fallback)"
    from lattice_qcd_rs_b200 import Context
    return Context


def test_magnetic_field_known_answers(backend):
    """field.rs:1580-1711, D = 3, 4^3, a = 1: cold lattice and single links set to i*1."""
    o = Oracle(3, 4, a=1.0)
    c = backend(3, 4, a=1.0)
    I3 = np.eye(3)
    X, Y = Oracle.sdir(0), Oracle.sdir(1)
    U = o.cold_links()
    c.links_upload(U)
    assert np.allclose(to_c(c.clover(X, Y)), 4 * I3, atol=EPS)          # every site of a cold lattice
    assert np.allclose(to_c(c.f_mu_nu(0, 1)), 0, atol=EPS)
    for d in range(3):
        assert np.allclose(to_c(c.magnetic_field(d)), 0, atol=EPS)
    U[0] = from_c(1j * I3)[0]                                             # link (origin, x) = i
    c.links_upload(U)
    assert np.allclose(to_c(c.clover(X, Y))[0], 2 * I3, atol=EPS)
    assert np.allclose(to_c(c.clover(Y, X))[0], 2 * I3, atol=EPS)
    assert np.allclose(to_c(c.f_mu_nu(0, 1))[0], 0, atol=EPS)
    U = o.cold_links()
    U[o.site_index([1, 0, 0]) * 3 + 1] = from_c(1j * I3)[0]               # link ((1,0,0), y) = i
    c.links_upload(U)
    assert np.allclose(to_c(c.clover(X, Y))[0], (3 + 1j) * I3, atol=EPS)
    assert np.allclose(to_c(c.clover(Y, X))[0], (3 - 1j) * I3, atol=EPS)
    assert np.allclose(to_c(c.f_mu_nu(0, 1))[0], 0.25j * I3, atol=EPS)
    assert np.allclose(to_c(c.magnetic_field(0))[0], 0, atol=EPS)
    assert np.allclose(to_c(c.magnetic_field(1))[0], 0, atol=EPS)
    assert np.allclose(to_c(c.magnetic_field(2))[0], 0.25 * I3, atol=EPS)


@pytest.mark.parametrize("D,ext", [(4, [4, 4, 4, 4]), (3, [4, 6, 2]), (2, [6, 4])])
def test_field_strength_matches_oracle(backend, D, ext):
    o = Oracle(D, ext, a=0.7)
    c = backend(D, ext, a=0.7)
    U = o.links_random(SEED_RNG, 3)
    U += 0.01 * np.sin(np.arange(U.size)).reshape(U.shape)  # off SU(3): no unitarity shortcut can hide
    c.links_upload(U)
    sites = range(0, o.ns, max(o.ns // 37, 1))
    for si, sj in ((1, 2), (-1, 2), (2, -1), (-D, -1), (D, 1)):
        got = to_c(c.clover(si, sj))
        for x in sites:
            want = o.clover(U, x, si, sj)
            assert np.abs(got[x] - want).max() <= 1e-12 * max(np.abs(want).max(), 1.0)
    got = to_c(c.f_mu_nu(0, D - 1))
    for x in sites:
        assert np.abs(got[x] - o.f_mu_nu(U, x, 1, D)).max() <= 1e-12
    if D >= 3:
        for d in range(D):
            got = to_c(c.magnetic_field(d))
            for x in sites:
                assert np.abs(got[x] - o.magnetic_field(U, x, d)).max() <= 1e-12
    from lattice_qcd_rs_b200 import LqError
    with pytest.raises(LqError):
        c.clover(0, 1)
    with pytest.raises(LqError):
        c.f_mu_nu(0, D)
