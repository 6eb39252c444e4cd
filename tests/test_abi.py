"""The C-ABI shared library loads (no GPU needed for that) and exports every symbol include/lqcd_b200.h declares;
the ctypes binding and the authored Rust shim name only symbols that exist; without a device the product library
refuses to create a state (there is no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "lqcd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lq_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from lattice_qcd_rs_b200 import _capi, build
    build.build()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 60
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/lqcd_b200.h but not exported"
    assert sorted(_capi.SYMBOLS) == declared, set(_capi.SYMBOLS) ^ set(declared)
    _capi.load()  # binds every symbol with its argument types


def test_python_constants_match_the_header_enums():
    """LQ_FLAG_*, LQ_OR_*, integrator and error-code values in include/lqcd_b200.h are the ones the ctypes binding uses."""
    from lattice_qcd_rs_b200 import _capi
    src = open(os.path.join(ROOT, "include", "lqcd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    enums = {k: int(v) for k, v in re.findall(r"\b(LQ_[A-Z0-9_]+)\s*=\s*(-?\d+)", src)}
    assert enums["LQ_FLAG_FOLD_HALO_SYNC"] == 512
    for name, value in enums.items():
        if name.startswith("LQ_FLAG_"):
            py = getattr(_capi, name[3:], None)
            assert py == value, (name, value, py)
        elif name.startswith("LQ_OR_"):
            assert getattr(_capi, name[3:]) == value, name
    assert _capi.INTEGRATOR_OMELYAN == enums["LQ_INTEGRATOR_OMELYAN"]


def test_rust_shim_binds_only_declared_symbols():
    ffi = open(os.path.join(ROOT, "rust", "lattice-qcd-b200", "src", "ffi.rs")).read()
    used = set(re.findall(r"pub fn (lq_[a-z0-9_]+)", ffi))
    assert used and used <= set(_header_symbols()), used - set(_header_symbols())


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is visible")
    from lattice_qcd_rs_b200 import Context, LqError
    with pytest.raises(LqError) as e:
        Context(4, 4)
    assert e.value.name == "LQ_E_NODEVICE"


def test_emu_library_is_not_reachable_from_the_package():
    """The host-emulation build is test infrastructure: nothing under lattice_qcd_rs_b200/ refers to it or to oracle/."""
    pkg = os.path.join(ROOT, "lattice_qcd_rs_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            text = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in text and "from oracle" not in text and "from tests" not in text, fn
