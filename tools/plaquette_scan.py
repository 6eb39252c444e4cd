"""Average plaquette of the SU(3) Wilson action from the textbook switches (sigma_3 fixed, coupling beta k/CA, sphere-
uniform direction): 1 heat-bath + 2 SU(2)-sub-group over-relaxation sweeps per iteration on L^4, against literature
values (e.g. Bali & Schilling 1993; Necco & Sommer 2002 tables): beta 5.7: 0.54920, 6.0: 0.59368, 6.2: 0.61363, 6.4: 0.63064.
  python tools/plaquette_scan.py [L=16] [iterations=400]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lattice_qcd_rs_b200 import Context, FLAG_PAULI3_FIXED, FLAG_UNIFORM_DIRECTION, OR_SU2_SUBGROUPS  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 400
LIT = {5.7: 0.54920, 6.0: 0.59368, 6.2: 0.61363, 6.4: 0.63064}
SEED = 0x457893F44AB067F0
for beta, lit in LIT.items():
    c = Context(4, L, a=1.0, beta=beta)
    c.set_flags(FLAG_PAULI3_FIXED | FLAG_UNIFORM_DIRECTION)
    c.links_set_random(SEED, 0)
    t = time.time()
    vals = []
    for k in range(N):
        c.sweep_heatbath(SEED, 1 + k, coupling_scale=1.0 / 3.0)
        c.sweep_overrelax(OR_SU2_SUBGROUPS)
        c.sweep_overrelax(OR_SU2_SUBGROUPS)
        if k >= N // 4:
            vals.append(c.average_trace_plaquette().real / 3.0)
    v = np.array(vals)
    nb = 10
    b = v[:len(v) // nb * nb].reshape(nb, -1).mean(1)
    err = b.std(ddof=1) / np.sqrt(nb)
    print(f"beta={beta}: <P>/3 = {v.mean():.5f} +- {err:.5f} (binned)   literature {lit:.5f}   diff {v.mean() - lit:+.5f}"
          f"   [{L}^4, {N} iterations, {time.time() - t:.1f} s]", flush=True)
