"""Wire format of the states: what serde_json / bincode 1.x write for the reference's
serde-derived structs (state.rs:654-659, 1047-1062; field.rs:28-31, 583-586, 1024-1027; lattice.rs:43-49).

The reference holds no serialized fixture and rustc is absent, so the format is pinned here by byte strings written
out BY HAND from the published serde rules of the crates involved (struct = fields in declaration order; bincode:
little endian, u64 lengths; nalgebra ArrayStorage = sequence of R*C column-major elements; Complex = (re, im)) --
"parity unpinned" against a real Rust binary, as lattice_qcd_rs_b200/serde_io.py states.
"""
import json
import struct

import numpy as np
import pytest

from lattice_qcd_rs_b200 import serde_io
from lattice_qcd_rs_b200 import state as lq


@pytest.fixture(params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def lib(request):
    if request.param == "emu":
        from tests import emu
        return emu.lib()
    import torch
    assert torch.cuda.is_available(), "the gpu-marked tests need a CUDA device (no CPU fallback)"
    return None


def _identity18():
    m = np.zeros((9, 2))
    m[0, 0] = m[4, 0] = m[8, 0] = 1.0
    return m.reshape(18)


def test_json_known_answer_cold_2x2():
    """D = 2, 2 points per direction, cold links: 8 identity matrices, written out literally."""
    links = np.tile(_identity18(), (8, 1))
    one = "[[1.0,0.0],[0.0,0.0],[0.0,0.0],[0.0,0.0],[1.0,0.0],[0.0,0.0],[0.0,0.0],[0.0,0.0],[1.0,0.0]]"
    expect = ('{"lattice":{"size":1.5,"dim":2},"beta":6.0,"link_matrix":{"data":[' + ",".join([one] * 8) + "]}}")
    text = serde_io.dumps_json(serde_io.state_to_json_obj(1.5, 2, 2, 6.0, links))
    assert text == expect
    back = serde_io.state_from_json_obj(json.loads(expect), 2)
    assert back["size"] == 1.5 and back["dim"] == 2 and back["beta"] == 6.0 and np.array_equal(back["links"], links)
    # with an E-field: e_field, t, lattice_state in declaration order (state.rs:1059-1061)
    e = np.arange(4 * 2 * 8, dtype=np.float64).reshape(8, 8) / 4.0
    obj = serde_io.ef_state_to_json_obj(1.5, 2, 2, 6.0, links, e, 7)
    assert list(obj.keys()) == ["e_field", "t", "lattice_state"]
    assert obj["e_field"]["data"][1][0] == {"data": [4.0, 4.25, 4.5, 4.75, 5.0, 5.25, 5.5, 5.75]}  # site 1, direction 0
    assert serde_io.dumps_json(obj).startswith('{"e_field":{"data":[[{"data":[0.0,0.25,0.5,0.75,1.0,1.25,1.5,1.75]},{"data":[2.0,')
    back = serde_io.ef_state_from_json_obj(json.loads(serde_io.dumps_json(obj)), 2)
    assert back["t"] == 7 and np.array_equal(back["e_field"], e) and np.array_equal(back["links"], links)


def test_bincode_known_answer():
    """The byte stream assembled field by field with struct.pack, independent of the numpy record writer."""
    rng = np.random.default_rng(3)
    D, dim = 2, 2
    nl, ns = dim ** D * D, dim ** D
    links = rng.normal(size=(nl, 18))
    e = rng.normal(size=(nl, 8))
    exp = struct.pack("<d", 0.75) + struct.pack("<Q", dim) + struct.pack("<d", 2.5) + struct.pack("<Q", nl)
    for l in range(nl):
        exp += struct.pack("<Q", 9)                       # ArrayStorage<_, 3, 3>: serialize_seq(Some(9))
        for k in range(9):                                # column-major elements, Complex = (re, im)
            exp += struct.pack("<dd", links[l, 2 * k], links[l, 2 * k + 1])
    got = serde_io.state_to_bincode(0.75, dim, D, 2.5, links)
    assert got == exp and len(got) == 32 + nl * 152
    back = serde_io.state_from_bincode(got, D)
    assert back["size"] == 0.75 and back["beta"] == 2.5 and np.array_equal(back["links"], links)
    expe = struct.pack("<Q", ns)
    for s in range(ns):
        expe += struct.pack("<Q", D)                      # SVector<Su3Adjoint, D>: sequence of D structs
        for d in range(D):
            expe += struct.pack("<Q", 8) + struct.pack("<8d", *e[s * D + d])   # Su3Adjoint { data: Vector8 }
    expe += struct.pack("<Q", 12) + exp                   # t, then lattice_state
    gote = serde_io.ef_state_to_bincode(0.75, dim, D, 2.5, links, e, 12)
    assert gote == expe
    back = serde_io.ef_state_from_bincode(gote, D)
    assert back["t"] == 12 and np.array_equal(back["e_field"], e) and np.array_equal(back["links"], links)
    # tuple-style fixed arrays (no per-matrix length)
    raw = serde_io.ef_state_to_bincode(0.75, dim, D, 2.5, links, e, 12, seq_prefix=False)
    assert len(raw) == 8 + ns * D * 64 + 8 + 32 + nl * 144
    back = serde_io.ef_state_from_bincode(raw, D, seq_prefix=False)
    assert np.array_equal(back["e_field"], e) and np.array_equal(back["links"], links)


def test_malformed_inputs():
    links = np.tile(_identity18(), (8, 1))
    good = serde_io.state_to_bincode(1.0, 2, 2, 1.0, links)
    with pytest.raises(ValueError):
        serde_io.state_from_bincode(good[:-1], 2)               # truncated
    with pytest.raises(ValueError):
        serde_io.state_from_bincode(good + b"\0", 2)            # trailing bytes
    with pytest.raises(ValueError):
        serde_io.state_from_bincode(good, 3)                    # 8 links are not a D = 3 lattice
    with pytest.raises(ValueError):
        serde_io.state_from_bincode(good, 2, seq_prefix=False)  # length prefixes read as data
    with pytest.raises(ValueError):
        serde_io.state_to_bincode(1.0, 2, 2, 1.0, links[:-1])   # IncompatibleSize
    with pytest.raises(ValueError):
        serde_io.state_from_json_obj({"lattice": {"size": 1.0, "dim": 1}, "beta": 1.0, "link_matrix": {"data": []}}, 2)
    bad = serde_io.state_to_json_obj(1.0, 2, 2, 1.0, links)
    bad["link_matrix"]["data"].pop()
    with pytest.raises(ValueError):
        serde_io.state_from_json_obj(bad, 2)


@pytest.mark.parametrize("D,n", [(4, 4), (3, 4)])
def test_state_round_trip_and_resume(lib, D, n):
    """Checkpoint / resume: a state written after k Monte-Carlo steps, read back into a fresh device context and
    continued with the checkpointed host generator reproduces the uninterrupted run bit for bit."""
    rng = lq.Rng(11)
    st = lq.LatticeStateEFSyncDefault.new_determinist(1.0, 6.0, n, rng, D=D, lib=lib)
    hb = lq.HeatBathSweep(rng)
    st2 = lq.LatticeStateDefault.new(st.lattice(), 6.0, st.link_matrix(), lib=lib)
    for _ in range(2):
        st2 = st2.monte_carlo_step(hb)
    blob, text, ck = st2.to_bincode(), st2.to_json(), hb.rng().checkpoint()
    for _ in range(2):
        st2 = st2.monte_carlo_step(hb)
    for restored in (lq.LatticeStateDefault.from_bincode(blob, D=D, lib=lib), lq.LatticeStateDefault.from_json(text, D=D, lib=lib)):
        assert restored.lattice() == st.lattice() and restored.beta() == 6.0
        hb2 = lq.HeatBathSweep(lq.Rng.from_checkpoint(ck))
        for _ in range(2):
            restored = restored.monte_carlo_step(hb2)
        assert np.array_equal(restored.link_matrix(), st2.link_matrix())
    # state with E-field and step counter
    st = st.simulate_symplectic_n(lq.SymplecticEulerCuda(), 0.01, 3)
    for back in (lq.LatticeStateEFSyncDefault.from_bincode(st.to_bincode(), D=D, lib=lib),
                 lq.LatticeStateEFSyncDefault.from_json(st.to_json(), D=D, lib=lib)):
        assert back.t() == st.t() == 3
        assert np.array_equal(back.link_matrix(), st.link_matrix()) and np.array_equal(back.e_field(), st.e_field())
        assert back.hamiltonian_total() == st.hamiltonian_total()
    with pytest.raises(lq.StateInitializationError):
        lq.LatticeStateDefault.from_bincode(serde_io.state_to_bincode(0.0, 2, 2, 1.0, np.zeros((8, 18))), D=2, lib=lib)
