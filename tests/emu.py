"""Host-emulation build of the kernel bodies (TEST INFRASTRUCTURE for the CPU CI; never on the product path).

`g++ -x c++ -DLQ_HOST_EMU` compiles lattice_qcd_rs_b200/csrc/lq_capi.cu with every kernel functor executed in a
plain host loop and malloc standing in for device memory.  It lets the CPU-only suite check the HOST LOGIC of the
C ABI (sequencing of kernels, buffer swaps, halo pack/unpack, multi-rank plumbing over gloo) and the kernel bodies'
arithmetic against the oracle without a GPU.  The package (lattice_qcd_rs_b200) never loads this library: its
loader opens only the CUDA .so and raises when that is missing.  The `-m gpu` tests re-run the same comparisons
through the real CUDA library on the B200.
"""
import os
import subprocess

from lattice_qcd_rs_b200 import _capi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "lattice_qcd_rs_b200", "csrc")
OUT = os.path.join(HERE, "_emu")
LIB = os.path.join(OUT, "liblqcd_emu.so")
DEPS = ["lq_capi.cu", "lq_kernels.cuh", "lq_common.cuh", "lq_local.cuh"]

_lib = None


def build():
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in DEPS] + [os.path.join(ROOT, "include", "lqcd_b200.h"), __file__]
    if os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in srcs):
        return LIB
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O2", "-march=x86-64-v3", "-DLQ_HOST_EMU", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared",
           os.path.join(CSRC, "lq_capi.cu"), "-o", LIB]
    subprocess.run(cmd, check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = _capi.bind(build())
    return _lib


def context(D, extent, **kw):
    return _capi.Context(D, extent, lib=lib(), **kw)
