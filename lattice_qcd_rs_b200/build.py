"""Build recipe of the CUDA library (sm_100a only): python -m lattice_qcd_rs_b200.build

Output: lattice_qcd_rs_b200/liblqcd_b200.so (in-tree, git-ignored, travels to the GPU box with the snapshot).
nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblqcd_b200.so")
SRCS = ["lq_capi.cu"]
DEPS = ["lq_capi.cu", "lq_kernels.cuh", "lq_common.cuh", "lq_local.cuh", "lq_tuned.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared", "-DLQ_BUILD_CUDA=1"]


USE_TUNED = True  # flipped on once lq_tuned.cuh exports the launchers (lq_tuned_efield_step, ...)


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    files = [os.path.join(CSRC, f) for f in DEPS] + [os.path.join(ROOT, "include", "lqcd_b200.h"), __file__]
    return any(os.path.exists(f) and os.path.getmtime(f) > t for f in files)


def build(force=False, verbose=False):
    if not (force or needs_build()):
        return LIB
    flags = list(NVCC_FLAGS)
    if USE_TUNED:
        flags.append("-DLQ_HAVE_TUNED=1")
    if verbose:
        flags += ["-Xptxas", "-v"]
    cmd = [nvcc(), *flags, *[os.path.join(CSRC, s) for s in SRCS], "-o", LIB]
    subprocess.run(cmd, check=True, cwd=HERE)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
