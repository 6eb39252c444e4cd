"""Times generic-functor paths (D = 3 and the D = 4 generic fallbacks): python tools/time_generic.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lattice_qcd_rs_b200 import Context, FLAG_GENERIC_KERNELS  # noqa: E402


def t(fn, n=10):
    fn()
    c.sync()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    c.sync()
    return (time.perf_counter() - t0) * 1e3 / n


c = Context(3, 160, a=1.0, beta=6.0)  # D = 3, 4.1 M sites
c.links_set_random(1, 0)
c.momenta_refresh(1, 1, 0.05)
print("D=3 160^3: symplectic step %.3f ms, plaquette %.3f ms, heat-bath sweep %.3f ms, reunitarize %.3f ms, gauss step %.3f ms" % (
    t(lambda: c.symplectic_n(0.001, 1)), t(lambda: c.average_trace_plaquette()), t(lambda: c.sweep_heatbath(1, 2)),
    t(lambda: c.reunitarize()), t(lambda: c.gauss_project_step())))
del c
c = Context(4, 32, a=1.0, beta=6.0)
c.set_flags(FLAG_GENERIC_KERNELS)
c.links_set_random(1, 0)
c.momenta_refresh(1, 1, 0.05)
print("D=4 32^4 generic functors: symplectic step %.3f ms, plaquette %.3f ms, heat-bath sweep %.3f ms, reunitarize %.3f ms, gauss step %.3f ms" % (
    t(lambda: c.symplectic_n(0.001, 1)), t(lambda: c.average_trace_plaquette()), t(lambda: c.sweep_heatbath(1, 2)),
    t(lambda: c.reunitarize()), t(lambda: c.gauss_project_step())))
