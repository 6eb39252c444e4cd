"""Times the Gauss-projection loop on L^4 for every iteration form: python tools/time_gauss.py [L]
  flags 0 / 64 / 128: transported-field loop (thread per site rolled / unrolled / thread per link);
  32: two-pass kernels; 4: one-pass functor (backward neighbours recomputed)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lattice_qcd_rs_b200 import Context  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
c = Context(4, L, a=1.0, beta=6.0)
c.links_set_random(0x457893F44AB067F0, 0)
for flags, name in ((0, "transported field, 168 regs / 12 warps"), (64, "transported field, 128 regs / 16 warps"),
                    (128, "transported field, 96 regs / 20 warps"), (192, "transported field, 80 regs / 24 warps"),
                    (32, "two-pass kernels"), (4, "one-pass functor")):
    c.set_flags(flags)
    best = 1e9
    for rep in range(3):
        c.momenta_refresh(0x457893F44AB067F0, 1)
        c.sync()
        t = time.perf_counter()
        it = c.gauss_project()
        c.sync()
        best = min(best, time.perf_counter() - t)
    print(f"flags {flags:3d} {name:48s} {it:4d} iterations  {best * 1e3:8.2f} ms  {best * 1e3 / it:7.4f} ms/iteration")
