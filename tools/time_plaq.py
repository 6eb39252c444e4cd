"""Times the plaquette reduction (tuned D = 4 kernel vs the generic functor) on L^4: python tools/time_plaq.py [L]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lattice_qcd_rs_b200 import Context, FLAG_GENERIC_KERNELS  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
c = Context(4, L, a=1.0, beta=6.0)
c.links_set_random(0x457893F44AB067F0, 0)
for flags, name in ((0, "tuned lq_plaq4_kernel"), (FLAG_GENERIC_KERNELS, "generic KPlaquette")):
    c.set_flags(flags)
    for _ in range(3):
        p = c.average_trace_plaquette()
    c.profile_enable(True)
    for _ in range(20):
        p = c.average_trace_plaquette()
    n, ms = c.profile_get("plaquette")
    c.profile_enable(False)
    nl = 4 * L ** 4
    print(f"{name:28s} {ms / n:8.4f} ms/call (kernel + final + sync)  {144.0 * nl / (ms / n) / 1e6:8.1f} GB/s(alg)  <P>={p}")
