import sys, time; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lattice_qcd_rs_b200 import Context, FLAG_PAULI3_FIXED, FLAG_UNIFORM_DIRECTION, INTEGRATOR_OMELYAN, INTEGRATOR_SYMPLECTIC_EULER, OMELYAN_LAMBDA
def binned_err(v, nb=10):
    m = len(v)//nb; b = v[:m*nb].reshape(nb, m).mean(1); return b.std(ddof=1)/np.sqrt(nb)
def run(L, dt, n, ntraj, kind, use_exp, sigma_fix=True, project=False, therm=150):
    c = Context(4, L, a=1.0, beta=6.0)
    seed=0x457893F44AB067F0
    c.set_flags(FLAG_PAULI3_FIXED|FLAG_UNIFORM_DIRECTION)
    c.links_set_random(seed,0)
    for k in range(therm): c.sweep_heatbath(seed,1+k,coupling_scale=1/3)
    p0 = c.average_trace_plaquette().real/3
    c.set_integrator(kind, OMELYAN_LAMBDA, use_exp)
    vals=[];acc=0;dh=[]
    t=time.time()
    for k in range(ntraj):
        r=c.hmc_trajectory(dt,n,seed,1000+k,sigma=(1/np.sqrt(6.0) if sigma_fix else 0.5/6.0),do_project=project)
        acc+=r["accepted"]; dh.append(r["h_new"]-r["h_old"])
        vals.append(c.average_trace_plaquette().real/3)
        if not use_exp and k % 10 == 9: c.reunitarize()
    v=np.array(vals[ntraj//5:])
    print(f"L={L} kind={kind} exp={use_exp} sigmafix={sigma_fix} proj={project} dt={dt} n={n}: start {p0:.5f} acc={acc/ntraj:.2f} <exp(-dH)>={np.mean(np.exp(-np.array(dh))):.3f} P={v.mean():.5f} +- {binned_err(v):.5f} (binned)  {time.time()-t:.1f}s", flush=True)
run(8, 0.1, 10, 2000, INTEGRATOR_OMELYAN, True)
run(8, 0.1, 20, 1000, INTEGRATOR_OMELYAN, True)
run(16, 0.08, 25, 500, INTEGRATOR_OMELYAN, True)
run(8, 0.05, 20, 1000, INTEGRATOR_SYMPLECTIC_EULER, True)
run(8, 0.02, 50, 600, INTEGRATOR_SYMPLECTIC_EULER, False)
try:
    run(8, 0.02, 50, 300, INTEGRATOR_SYMPLECTIC_EULER, False, sigma_fix=False, project=True)
except Exception as e:  # the reference's own recipe (sigma = 0.5/beta, Gauss projection, Euler links)
    print("reference recipe:", e)
