"""Multi-rank path: domain decomposition + halo exchange + global sums.

CPU CI: world_size 2 and 4 over gloo with the host-emulation build of the kernels (tests/emu.py); the decomposed
run must reproduce the single-rank oracle on the same global lattice.  The gpu-marked variant launches the same
worker over NCCL when the box has >= 2 GPUs.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 0x457893F44AB067F0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, backend, D, gext, proc_grid, q, transport="p2p", flags=0):
    try:
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        os.environ.setdefault("OMP_NUM_THREADS", "2")
        os.environ["LQ_HALO_TRANSPORT"] = transport
        if backend == "nccl":
            torch.cuda.set_device(rank)
        dist.init_process_group(backend, rank=rank, world_size=world)
        from lattice_qcd_rs_b200.dist import DistContext
        from oracle.oracle import Oracle
        lib = None
        if backend == "gloo":
            from tests import emu
            lib = emu.lib()
        o = Oracle(D, gext, a=1.0, beta=6.0)
        dc = DistContext(D, gext, a=1.0, beta=6.0, proc_grid=proc_grid, lib=lib)
        c = dc.ctx
        if flags:
            c.set_flags(flags)
        U = o.links_random(SEED)
        E = o.momenta_refresh(SEED, 5)
        res = {}

        def rel(a, b):
            return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

        c.links_upload(dc.scatter(U, 18))
        c.efield_upload(dc.scatter(E, 8))
        res["roundtrip"] = float(np.abs(dc.gather(c.links_download(), 18) - U).max())
        res["plaq"] = abs(c.plaquette_sum() - o.plaquette_sum(U)) / abs(o.plaquette_sum(U))
        res["h"] = abs(c.hamiltonian_total() - o.hamiltonian_total(U, E)) / o.hamiltonian_total(U, E)
        res["force"] = rel(dc.gather(c.force(), 8), o.force(U))
        res["staples"] = rel(dc.gather(c.staples(), 18), o.staples(U))
        res["gauss"] = rel(dc.gather(c.gauss_field(), 18), o.gauss_field(U, E))
        res["gauss_div"] = abs(c.gauss_sum_div() - o.gauss_sum_div(U, E)) / o.gauss_sum_div(U, E)
        c.gauss_project_step()
        res["gauss_step"] = rel(dc.gather(c.efield_download(), 8), o.project_to_gauss_step(U, E))
        # MD trajectory
        c.efield_upload(dc.scatter(E, 8))
        c.symplectic_n(0.01, 4)
        Uo, Eo = o.integrate(U, E, "symplectic", 0.01, n=4)
        res["md_U"] = rel(dc.gather(c.links_download(), 18), Uo)
        res["md_E"] = rel(dc.gather(c.efield_download(), 8), Eo)
        for kind, name in enumerate(["sync_sync", "leap_leap", "sync_leap", "leap_sync"]):
            c.links_upload(dc.scatter(U, 18))
            c.efield_upload(dc.scatter(E, 8))
            c.integrate(kind, 0.01)
            Uo, Eo = o.integrate(U, E, name, 0.01)
            res["int_" + name] = max(rel(dc.gather(c.links_download(), 18), Uo),
                                     rel(dc.gather(c.efield_download(), 8), Eo))
        # integrator options (lq_set_integrator / lq_md_n): Omelyan steps with the exponential link update, every
        # kick fused with the link step that follows it (the push variant of the fused kernel on decomposed contexts)
        c.links_upload(dc.scatter(U, 18))
        c.efield_upload(dc.scatter(E, 8))
        c.set_integrator(1, 0.1931833275037836, True)
        c.md_n(0.02, 3)
        c.set_integrator()
        Uo, Eo = o.md_n(U, E, 0.02, 3, kind=1, use_exp=True)
        res["omelyan_U"] = rel(dc.gather(c.links_download(), 18), Uo)
        res["omelyan_E"] = rel(dc.gather(c.efield_download(), 8), Eo)
        # full HMC trajectory with Philox momenta: decomposition-independent streams, same accept decision
        c.links_upload(dc.scatter(U, 18))
        r = c.hmc_trajectory(0.01, 5, SEED, 3)
        ro = o.hmc_trajectory(U, 0.01, 5, SEED, 3)
        res["hmc_h"] = abs(r["h_new"] - ro["h_new"]) / abs(ro["h_new"])
        res["hmc_acc"] = float(r["accepted"] != ro["accepted"]) + float(r["gauss_steps"] != ro["gauss_steps"])
        res["hmc_U"] = rel(dc.gather(c.links_download(), 18), ro["U"])
        # sweeps
        c.links_upload(dc.scatter(U, 18))
        c.sweep_heatbath(SEED, 11)
        res["heatbath"] = rel(dc.gather(c.links_download(), 18), o.sweep_heatbath(U, SEED, 11))
        c.links_upload(dc.scatter(U, 18))
        c.sweep_overrelax(0)
        res["overrelax"] = rel(dc.gather(c.links_download(), 18), o.sweep_overrelax(U, 0))
        c.links_upload(dc.scatter(U, 18))
        na, sp = c.sweep_metropolis(SEED, 13, spread=0.1, n_update=2)
        Uo, nao, spo = o.sweep_metropolis(U, SEED, 13, n_update=2, spread=0.1)
        res["metropolis"] = rel(dc.gather(c.links_download(), 18), Uo) + abs(na - nao) + abs(sp - spo) / spo
        res["exchanges"] = dc.halo_exchanges + c.p2p_exchanges
        res["transport"] = dc.transport
        if rank == 0:
            q.put(res)
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback
        traceback.print_exc()
        if rank == 0:
            q.put({"error": repr(e)})
        raise


def _run(world, backend, D, gext, proc_grid, transport="p2p", flags=0):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, backend, D, gext, proc_grid, q, transport, flags))
             for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
    assert "error" not in res, res
    return res


FOLD_ALL = 512 | 2048  # LQ_FLAG_FOLD_HALO_SYNC for the projection loop and (| 2048) the MD chain
TOL = {"roundtrip": 0.0, "hmc_acc": 0.0}


def _check(res):
    assert res["exchanges"] > 0
    for k, v in res.items():
        if k in ("exchanges", "transport"):
            continue
        tol = TOL.get(k, 1e-9 if k in ("heatbath", "overrelax", "metropolis", "hmc_U") else 1e-12)
        assert v <= tol, (k, v, res)


@pytest.mark.parametrize("world,D,gext,proc_grid", [
    (2, 4, [4, 4, 4, 8], [1, 1, 1, 2]),
    (4, 4, [4, 4, 4, 8], [1, 1, 2, 2]),
    (2, 3, [4, 6, 4], [1, 1, 2]),
    (8, 4, [4, 4, 4, 8], [1, 1, 2, 4]),  # the process grid of the 8-GPU bench (z x t = 2 x 4, with corners)
])
def test_decomposed_matches_oracle_gloo(world, D, gext, proc_grid):
    _check(_run(world, "gloo", D, gext, proc_grid))


@pytest.mark.gpu
def test_decomposed_matches_oracle_nccl():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    for transport in ("p2p", "nccl"):  # peer-memory pushes by our own kernels / NCCL send-recv through the callbacks
        res = _run(2, "nccl", 4, [8, 8, 8, 16], [1, 1, 1, 2], transport)
        _check(res)
        assert res["transport"] == ("p2p" if transport == "p2p" else "nccl-callbacks"), res["transport"]
        if transport == "p2p":  # x0 extent 32: the bulk-copy (TMA) AoS <-> SoA row kernels on a lattice with ghost layers
            _check(_run(2, "nccl", 4, [32, 4, 4, 4], [1, 1, 1, 2], transport))
            # opt-in: halo synchronisation folded into the MD and projection kernels (LQ_FLAG_FOLD_HALO_SYNC | 2048)
            _check(_run(2, "nccl", 4, [32, 4, 4, 4], [1, 1, 1, 2], transport, flags=FOLD_ALL))
        if n >= 4:
            # 16 x 8 sites per (z, t) column: the launches with the halo synchronisation folded in cover this geometry
            res = _run(4, "nccl", 4, [16, 8, 8, 8], [1, 1, 2, 2], transport)
            _check(res)
            if transport == "p2p":
                _check(_run(4, "nccl", 4, [16, 8, 8, 8], [1, 1, 2, 2], transport, flags=FOLD_ALL))
            assert res["transport"] == ("p2p" if transport == "p2p" else "nccl-callbacks"), res["transport"]
        if n >= 8 and transport == "p2p":  # the 2(z) x 4(t) grid the 8-GPU bench runs on, incl. the zt corners
            res = _run(8, "nccl", 4, [16, 8, 8, 16], [1, 1, 2, 4], transport)
            _check(res)
            assert res["transport"] == "p2p", res["transport"]
