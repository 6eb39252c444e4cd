"""Times the two kernels of a Gauss projection iteration on 32^4 (CUDA events around every launch): python tools/time_gauss.py"""
import sys, os
sys.path.insert(0, ".")
from lattice_qcd_rs_b200 import Context
c = Context(4, 32, a=1.0, beta=6.0)
c.links_set_random(1, 0); c.momenta_refresh(1, 1, 0.08)
for _ in range(8): c.gauss_project_step()
c.profile_enable(True)
for _ in range(60): c.gauss_project_step()
n1, m1 = c.profile_get("gauss_field"); n2, m2 = c.profile_get("gauss_step")
print("GF %.4f ms  STEP %.4f ms" % (m1 / n1, m2 / n2))
