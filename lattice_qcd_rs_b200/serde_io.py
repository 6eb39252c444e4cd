"""On-disk / wire format of the states: what `serde` writes for the reference's own types.

The reference derives `Serialize`/`Deserialize` (feature `serde-serialize`, on by default, Cargo.toml:19-22) for

  LatticeCyclic<D>            { size: f64, dim: usize }                               lattice.rs:43-49
  LinkMatrix                  { data: Vec<Matrix3<Complex<f64>>> }                    field.rs:583-586
  Su3Adjoint                  { data: Vector8<f64> }                                  field.rs:28-31
  EField<D>                   { data: Vec<SVector<Su3Adjoint, D>> }                   field.rs:1024-1027
  LatticeStateDefault<D>      { lattice, beta: f64, link_matrix }                     state.rs:654-659
  LatticeStateEFSyncDefault   { e_field, t: usize, lattice_state }                    state.rs:1047-1062

and ships no format crate of its own, so the two de-facto formats are restated here: `serde_json` (structs ->
objects keyed by field name, sequences -> arrays) and `bincode` 1.x default options (little endian, fixed-width
integers, `usize` and sequence lengths as u64, struct fields concatenated in declaration order).

Third-party rules this relies on (un-vendored dependencies, restated from their published sources; PARITY UNPINNED:
there is no rustc here and the reference holds no serialized fixture to compare against):
  * nalgebra ^0.31 `Matrix<T,R,C,ArrayStorage>` serializes its storage, and `ArrayStorage<T,R,C>` calls
    `serialize_seq(Some(R*C))` over `as_slice()` -- a SEQUENCE of R*C elements in column-major order (so bincode
    writes a u64 length in front of every static matrix / vector; `seq_prefix=False` drops it for releases that
    serialize fixed arrays as tuples).  The C ABI's host layout (18 f64 per link, column-major) is that slice.
  * num-complex `Complex<T>` serializes as the tuple `(re, im)`.
JSON numbers are written with Python's shortest round-trip repr; serde_json (ryu) prints the same digits but a
different exponent style for some magnitudes (`1e-7` vs `1e-07`) -- both parse to the same f64 on either side.

All functions work on host arrays in the reference AoS layouts (what lq_links_download / lq_efield_download return);
`LatticeStateDefault.to_json()/to_bincode()` and the `from_*` constructors in state.py wrap them.
"""
import json
import struct

import numpy as np


# ------------------------------------------------------------------------------------------------ plain-data form
def _links_aos(links, n_links):
    a = np.ascontiguousarray(np.asarray(links, dtype=np.float64)).reshape(-1)
    if a.size != n_links * 18:
        raise ValueError(f"link array has {a.size} f64, expected {n_links * 18}")
    return a.reshape(n_links, 9, 2)


def _efield_aos(e_field, n_sites, D):
    a = np.ascontiguousarray(np.asarray(e_field, dtype=np.float64)).reshape(-1)
    if a.size != n_sites * D * 8:
        raise ValueError(f"E-field array has {a.size} f64, expected {n_sites * D * 8}")
    return a.reshape(n_sites, D, 8)


def _check_lattice(size, dim, D):
    # structural only: serde's derived Deserialize does not run LatticeCyclic::new's checks (lattice.rs:190-201);
    # state.py applies them when it builds a device state from the decoded fields
    if dim < 1 or D < 1:
        raise ValueError("invalid lattice (dim and D must be >= 1)")


# ------------------------------------------------------------------------------------------------ serde_json
def state_to_json_obj(size, dim, D, beta, links):
    """LatticeStateDefault<D> as the object serde_json writes (state.rs:654-659)."""
    m = _links_aos(links, dim ** D * D)
    return {"lattice": {"size": float(size), "dim": int(dim)}, "beta": float(beta), "link_matrix": {"data": m.tolist()}}


def ef_state_to_json_obj(size, dim, D, beta, links, e_field, t):
    """LatticeStateEFSyncDefault<LatticeStateDefault<D>, D> (state.rs:1047-1062): e_field, t, lattice_state."""
    e = _efield_aos(e_field, dim ** D, D)
    return {"e_field": {"data": [[{"data": comp} for comp in site] for site in e.tolist()]}, "t": int(t),
            "lattice_state": state_to_json_obj(size, dim, D, beta, links)}


def dumps_json(obj):
    return json.dumps(obj, separators=(",", ":"))  # serde_json::to_string: no whitespace, struct field order kept


def state_from_json_obj(obj, D):
    """-> dict(size, dim, D, beta, links (Nl, 18))."""
    lat = obj["lattice"]
    size, dim = float(lat["size"]), int(lat["dim"])
    _check_lattice(size, dim, D)
    data = np.asarray(obj["link_matrix"]["data"], dtype=np.float64)
    nl = dim ** D * D
    if data.shape != (nl, 9, 2):
        raise ValueError(f"link_matrix.data has shape {data.shape}, expected {(nl, 9, 2)} for D={D}, dim={dim}")
    return {"size": size, "dim": dim, "D": D, "beta": float(obj["beta"]), "links": data.reshape(nl, 18)}


def ef_state_from_json_obj(obj, D):
    st = state_from_json_obj(obj["lattice_state"], D)
    ns = st["dim"] ** D
    rows = obj["e_field"]["data"]
    e = np.asarray([[c["data"] for c in site] for site in rows], dtype=np.float64)
    if e.shape != (ns, D, 8):
        raise ValueError(f"e_field.data has shape {e.shape}, expected {(ns, D, 8)}")
    st.update(e_field=e.reshape(ns * D, 8), t=int(obj["t"]))
    return st


# ------------------------------------------------------------------------------------------------ bincode 1.x
def _u64(v):
    return struct.pack("<Q", int(v))


def _link_records(m, seq_prefix):
    nl = m.shape[0]
    if not seq_prefix:
        return m.astype("<f8").tobytes()
    rec = np.empty(nl, dtype=[("n", "<u8"), ("v", "<f8", (18,))])
    rec["n"] = 9
    rec["v"] = m.reshape(nl, 18)
    return rec.tobytes()


def state_to_bincode(size, dim, D, beta, links, seq_prefix=True):
    """bincode::serialize(&LatticeStateDefault<D>): f64 size, u64 dim, f64 beta, u64 Nl, Nl x [u64 9,] 9 x (re, im)."""
    m = _links_aos(links, dim ** D * D)
    return b"".join([struct.pack("<dQd", float(size), int(dim), float(beta)), _u64(m.shape[0]),
                     _link_records(m, seq_prefix)])


def ef_state_to_bincode(size, dim, D, beta, links, e_field, t, seq_prefix=True):
    """bincode::serialize(&LatticeStateEFSyncDefault<..>): u64 Ns, Ns x ([u64 D,] D x ([u64 8,] 8 x f64)), u64 t, state."""
    e = _efield_aos(e_field, dim ** D, D)
    ns = e.shape[0]
    if seq_prefix:
        rec = np.empty(ns, dtype=[("n", "<u8"), ("c", [("n", "<u8"), ("v", "<f8", (8,))], (D,))])
        rec["n"] = D
        rec["c"]["n"] = 8
        rec["c"]["v"] = e
        body = rec.tobytes()
    else:
        body = e.astype("<f8").tobytes()
    return b"".join([_u64(ns), body, _u64(t), state_to_bincode(size, dim, D, beta, links, seq_prefix)])


class _Reader:
    def __init__(self, buf):
        self.b, self.o = memoryview(buf), 0

    def take(self, n):
        if self.o + n > len(self.b):
            raise ValueError("bincode: unexpected end of input")
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def u64(self):
        return struct.unpack("<Q", self.take(8))[0]

    def f64(self):
        return struct.unpack("<d", self.take(8))[0]


def _read_state(r, D, seq_prefix):
    size, dim, beta = r.f64(), r.u64(), r.f64()
    _check_lattice(size, dim, D)
    nl = r.u64()
    if nl != dim ** D * D:
        raise ValueError(f"bincode: {nl} link matrices, expected {dim ** D * D} for D={D}, dim={dim}")
    if seq_prefix:
        rec = np.frombuffer(r.take(nl * 152), dtype=[("n", "<u8"), ("v", "<f8", (18,))])
        if nl and not np.all(rec["n"] == 9):
            raise ValueError("bincode: a link matrix does not carry the sequence length 9 (try seq_prefix=False)")
        links = np.array(rec["v"], dtype=np.float64)
    else:
        links = np.frombuffer(r.take(nl * 144), dtype="<f8").astype(np.float64).reshape(nl, 18)
    return {"size": size, "dim": int(dim), "D": D, "beta": beta, "links": links}


def state_from_bincode(buf, D, seq_prefix=True):
    r = _Reader(buf)
    st = _read_state(r, D, seq_prefix)
    if r.o != len(r.b):
        raise ValueError("bincode: trailing bytes")
    return st


def ef_state_from_bincode(buf, D, seq_prefix=True):
    r = _Reader(buf)
    ns = r.u64()
    if seq_prefix:
        dt = np.dtype([("n", "<u8"), ("c", [("n", "<u8"), ("v", "<f8", (8,))], (D,))])
        rec = np.frombuffer(r.take(ns * dt.itemsize), dtype=dt)
        if ns and not (np.all(rec["n"] == D) and np.all(rec["c"]["n"] == 8)):
            raise ValueError("bincode: E-field sequence lengths do not match D / 8 (try seq_prefix=False)")
        e = np.array(rec["c"]["v"], dtype=np.float64)
    else:
        e = np.frombuffer(r.take(ns * D * 64), dtype="<f8").astype(np.float64)
    t = r.u64()
    st = _read_state(r, D, seq_prefix)
    if r.o != len(r.b):
        raise ValueError("bincode: trailing bytes")
    if ns != st["dim"] ** D:
        raise ValueError(f"bincode: E-field has {ns} sites, lattice has {st['dim'] ** D}")
    st.update(e_field=e.reshape(ns * D, 8), t=int(t))
    return st
