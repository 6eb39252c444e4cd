"""Small, fixed workload for ncu: one HMC trajectory (+ one heat-bath and one over-relaxation sweep) on L^4.
  ncu ... python tools/prof_target.py [--extent 32] [--md-steps 10] [--sweeps]
Numbers printed by a run under ncu are never bench values."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lattice_qcd_rs_b200 import Context  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--extent", type=int, default=32)
ap.add_argument("--md-steps", type=int, default=10)
ap.add_argument("--sweeps", action="store_true")
ap.add_argument("--no-project", action="store_true")
a = ap.parse_args()
SEED = 0x457893F44AB067F0
c = Context(4, a.extent, a=1.0, beta=6.0)
c.links_set_random(SEED, 0)
r = c.hmc_trajectory(0.01, a.md_steps, SEED, 1, do_project=not a.no_project)
c.reunitarize()
if a.sweeps:
    c.sweep_heatbath(SEED, 2)
    c.sweep_overrelax(0)
    c.sweep_metropolis(SEED, 3)
print(r, c.average_trace_plaquette(), c.kernel_launches)
