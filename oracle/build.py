"""Build recipe for the CPU oracle (test infrastructure only; never on the product path).

The reference (/root/reference) is a Rust crate and this image has no rustc/cargo, so
`oracle/_ref` (the real reference compiled here) cannot exist: the oracle is a C++
restatement ("port") of the reference loops, see lqcd_oracle.hpp.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liblqcd_oracle.so")
SRCS = ["lqcd_oracle_capi.cpp"]
HDRS = ["lqcd_oracle.hpp"]
# -ffp-contract=off: Rust never fuses a*b+c, so the restatement must not either.
# x86-64-v3 (not -march=native): the .so is built in the CPU container and travels to the GPU box.
FLAGS = ["-O3", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-std=c++17",
         "-fno-math-errno", "-Wall", "-Wno-unknown-pragmas"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SRCS + HDRS + ["build.py"])


def build(force: bool = False) -> str:
    if force or needs_build():
        cmd = ["g++", *FLAGS, *[os.path.join(HERE, s) for s in SRCS], "-o", LIB]
        subprocess.run(cmd, check=True, cwd=HERE)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
