"""Times the MD loop variants on L^4 (per MD step): python tools/time_md.py [L]
  reference symplectic Euler (lq_symplectic_n) | leap-frog + exponential update | Omelyan (+ exp)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lattice_qcd_rs_b200 import (Context, INTEGRATOR_OMELYAN, INTEGRATOR_SYMPLECTIC_EULER,  # noqa: E402
                                 OMELYAN_LAMBDA)

L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
c = Context(4, L, a=1.0, beta=6.0)
c.links_set_random(0x457893F44AB067F0, 0)
c.momenta_refresh(1, 1, 0.05)
n = 40
for name, kind, ex in (("symplectic Euler (reference)", INTEGRATOR_SYMPLECTIC_EULER, False),
                       ("leap-frog + exp", INTEGRATOR_SYMPLECTIC_EULER, True),
                       ("Omelyan + Euler", INTEGRATOR_OMELYAN, False), ("Omelyan + exp", INTEGRATOR_OMELYAN, True)):
    c.set_integrator(kind, OMELYAN_LAMBDA, ex)
    c.md_n(0.001, 4)
    c.sync()
    t = time.perf_counter()
    c.md_n(0.001, n)
    c.sync()
    ms = (time.perf_counter() - t) * 1e3 / n
    print(f"{name:30s} {ms:7.3f} ms per MD step   {4 * L ** 4 / ms / 1e6:7.2f} G link-updates/s")
