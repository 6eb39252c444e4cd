#!/usr/bin/env python
"""bench.py -- HMC link-updates/sec at 32^4, f64, on N B200s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config metric|c4|c5|d3]
  (N > 1: launched by torch.distributed.run, one rank per GPU; a CUDA-vs-oracle parity block on a small lattice over the
   same process grid runs before timing and is printed as "parity"; "per_rank" lists every rank's kernel times)
  --flags: LQ_FLAG_* bits for A/B runs (512 | 2048: halo synchronisation folded into the kernels; 32: two-pass Gauss loop)

A "step" is one HMC trajectory of HybridMonteCarloDiagnostic::next_element (hybrid_monte_carlo.rs:465-471, 573-613):
momentum refresh -> Gauss projection -> H_old -> 100 symplectic-Euler MD steps (dt = 0.01, the reference's own HMC
bench shape, benches/bench.rs:166-189) -> H_new -> accept/reject, followed by normalize_link_matrices.  One link-update
= one link carried through one MD step, so a step is 100 * N_links link-updates and ALL the per-trajectory overhead is
inside the timed region.

PINNED WORKLOAD: every timed step starts from the same device snapshot of the hot-start lattice (lq_restore: one
device-to-device copy inside the timed region) and uses the same momentum stream, so every step is the SAME trajectory
-- same Gauss-projection iteration count (data dependent: printed), same accept decision -- and ms_per_step does not
depend on --steps / --warmup.  (The crate's recipe -- sigma = 0.5/beta, Euler link update -- does not equilibrate: a
free-running chain drifts and its Gauss iteration count climbs with the trajectory index)

--config metric (default): 32^4 per GPU, weak scaling (split along t, and z at 8).   value / e2e as below.
--config c4: BASELINE config 4, ONE 48^3 x 96 lattice strong-scaled over the N GPUs (t split; z too at 8).
--config c5: BASELINE config 5, ONE 64^4 lattice over the N GPUs (2 x 32^4 per GPU at N = 8).
--config d3: BASELINE config 5b, the dimension-generic API on a D = 3 40^3 lattice (single GPU).

value : device-resident (links stay in HBM between trajectories), CUDA events on the context stream, max over ranks.
e2e   : the same trajectory through the C ABI with HOST buffers: every step uploads its links from and downloads its result
        to pinned host memory inside the timed region, through the pipelined marshalling calls (the upload of step k+1 and
        the download of step k travel while step k computes); e2e.serial_calls: the synchronous upload -> trajectory ->
        download sequence; e2e.host_copies_alone: the copies without any compute.
--impl reference : the CPU restatement of the reference loops (oracle/, OpenMP over all host cores; the Rust crate
        cannot be built in this image) on a bounded sample of the same 32^4 workload (see CpuSample).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 0x457893F44AB067F0
BETA, SPACING, DT, MD_STEPS = 6.0, 1.0, 0.01, 100
TRAJ_COUNTER = 1  # momentum stream of THE trajectory every step repeats
# Gauss-projection iterations of that trajectory on the 32^4 hot start (data dependent; bench.py prints the measured
# count next to this constant, and the CPU arm derives its MD : Gauss mix from it)
GAUSS_STEPS_PINNED = 101
# algorithmic bytes per link of the dominant kernel (fused force + E kick + link step), DESIGN.md section 4:
# read U 144 + read E 64 + write E 64 + write U' 144
BYTES_FUSED = 416
# its f64 flops per link: 12 staple matmuls + U*A (13 x 216) + trace/kick (~60) + E.to_matrix * U + Euler (~280)
FLOPS_FUSED = 13 * 216 + 60 + 280


def workload_text(config, L):
    """The workload both arms name in config.workload (the reference arm runs a bounded sample of it: cpu_baseline.sample)."""
    name = {"metric": f"{L}^4 per GPU", "c4": "BASELINE config 4: ONE 48^3 x 96 lattice over all GPUs",
            "c5": "BASELINE config 5: ONE 64^4 lattice over all GPUs",
            "d3": "BASELINE config 5b: D = 3, 40^3, dimension-generic kernels"}[config]
    return (f"{name}, beta={BETA}, a={SPACING}: full HMC trajectory (refresh sigma=0.5/beta, Gauss projection, 2x H_total, "
            f"{MD_STEPS} symplectic-Euler steps dt={DT}, accept/reject) + normalize_link_matrices; hot start (Philox random "
            "SU(3)); every step repeats the SAME trajectory from a device snapshot (pinned workload)")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        sm, mx, pw, reasons = [], [], [], set()
        with open(self.tmp.name) as f:
            for line in f:
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1]))
                    mx.append(float(c[2]))
                    pw.append(float(c[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.tmp.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=float(max(pw)))
        return out


# The CPU arm runs a BOUNDED SAMPLE of the same 32^4 workload: a full trajectory costs minutes on the host cores, so a
# sample step is 1/25 of one -- SAMPLE_MD symplectic steps and round(SAMPLE_MD * G / 100) Gauss-projection iterations,
# G = the Gauss iteration count of the pinned GPU trajectory -- on the full 32^4 hot lattice.  What a sample leaves out
# (momentum refresh, 2 x H_total, accept, normalise) is < 1 % of a trajectory on either side.
SAMPLE_MD = 4


def sample_gauss(md, gauss_per_traj=GAUSS_STEPS_PINNED):
    return max(1, int(round(md * gauss_per_traj / float(MD_STEPS))))


class CpuSample:
    def __init__(self, ext, md=SAMPLE_MD, gauss=None, literal=True):
        from oracle.oracle import Oracle
        self.o = o = Oracle(4, ext, a=SPACING, beta=BETA)
        o.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1; this arm uses every host core
        self.U = o.links_random(SEED)
        self.E = o.momenta_refresh(SEED, 1)
        self.ext, self.md, self.gauss, self.literal = ext, md, sample_gauss(md) if gauss is None else gauss, literal

    def step(self):
        o = self.o
        for _ in range(self.gauss):
            self.E = o.project_to_gauss_step(self.U, self.E)
        self.U, self.E = o.integrate(self.U, self.E, "symplectic", DT, n=self.md, literal=self.literal)

    def describe(self):
        mode = ("literal reference loops incl. the 28-matmul derivative_e (state.rs:1439-1441)" if self.literal else
                "optimised CPU loops (13-matmul derivative_e: U * A formed once, then the eight traces)")
        return (f"each step = {self.md} symplectic-Euler MD steps (dt={DT}) + {self.gauss} Gauss-projection iterations "
                f"on the full {self.ext}^4 beta={BETA} hot lattice (the pinned GPU trajectory's {MD_STEPS} : "
                f"{GAUSS_STEPS_PINNED} mix scaled down; momentum refresh, 2x H_total, accept and normalise left out: "
                f"< 1 % of a trajectory), {mode}, g++ -O3 -fopenmp on all host cores; C++ restatement of "
                "lattice-qcd-rs v0.2.1 (oracle/), not the Rust binary")


def cpu_sample(ext, steps=1, warmup=0, budget_s=240.0, literal=True, gauss=None):
    """Oracle (C++/OpenMP restatement of the reference loops) timed on the host cores.  The sample per step shrinks
    (4 -> 2 -> 1 MD steps, Gauss iterations in proportion) when steps x (time of one step) would exceed `budget_s`."""
    warmup = min(warmup, 1)  # no clock ramp or JIT on the CPU side: one step faults the pages in
    cs = CpuSample(ext, literal=literal, gauss=gauss)
    t0 = time.perf_counter()
    cs.step()  # calibration (counts as the warm-up step)
    t1 = time.perf_counter() - t0
    md0, g0 = cs.md, cs.gauss
    for md in (2, 1):
        if t1 * cs.md / md0 * steps > budget_s:
            cs.md, cs.gauss = md, max(1, int(round(g0 * md / md0)))
    if warmup == 0 and steps == 1 and cs.md == md0:
        el, done = t1, 1  # the calibration step IS the sample (default bench.py run: about 10 s of CPU work)
    else:
        t0 = time.perf_counter()
        for _ in range(steps):
            cs.step()
        el, done = time.perf_counter() - t0, steps
    val = cs.md * cs.o.nl * done / el
    return dict(value=val, unit="link-updates/s", cores=cs.o.num_threads(), kind="port",
                sample=cs.describe() + f"; {done} step(s), {el:.1f} s"), el


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu, el = cpu_sample(args.extent, steps=args.steps, warmup=args.warmup)
    opt, _ = cpu_sample(args.extent, steps=1, warmup=0, literal=False)  # the "optimised CPU" figure beside it
    cpu["optimised_cpu"] = {"value": opt["value"], "unit": opt["unit"], "sample": opt["sample"]}
    val = cpu["value"]
    line = {
        "impl": "reference", "metric": "HMC link-updates/sec at 32^4 f64", "value": val, "unit": "link-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text("metric", args.extent), "config": "metric", "sample": cpu["sample"]},
        "cpu_baseline": cpu,
        "e2e": {"value": val, "unit": "link-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- multi-GPU parity
def dist_parity(world, pg):
    """CUDA-vs-oracle comparison on a SMALL global lattice decomposed over the same process grid as the timed run, in
    the same process group, before timing (oracle = checker only).  Every rank computes the oracle (seconds)."""
    import torch.distributed as dist
    from lattice_qcd_rs_b200.dist import DistContext
    from oracle.oracle import Oracle
    gext = [16, 8, 8, 8]  # x0 * x1 = 128 sites per (z, t) column: the kernels with the halo synchronisation folded in apply
    for d in range(4):
        gext[d] = max(gext[d], 4 * pg[d])
    o = Oracle(4, gext, a=SPACING, beta=BETA)
    o.set_num_threads(max(1, (os.cpu_count() or 1) // world))
    dc = DistContext(4, gext, a=SPACING, beta=BETA, proc_grid=pg)
    c = dc.ctx
    U = o.links_random(SEED)
    E = o.momenta_refresh(SEED, 5)

    def rel(a, b):
        return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

    errs = {}
    c.links_upload(dc.scatter(U, 18))
    c.efield_upload(dc.scatter(E, 8))
    errs["plaquette"] = abs(c.plaquette_sum() - o.plaquette_sum(U)) / abs(o.plaquette_sum(U))
    errs["force"] = rel(dc.gather(c.force(), 8), o.force(U))
    c.symplectic_n(DT, 3)
    Uo, Eo = o.integrate(U, E, "symplectic", DT, n=3)
    errs["md_links"] = rel(dc.gather(c.links_download(), 18), Uo)
    errs["md_efield"] = rel(dc.gather(c.efield_download(), 8), Eo)
    c.links_upload(dc.scatter(U, 18))
    r = c.hmc_trajectory(DT, 5, SEED, 3)
    ro = o.hmc_trajectory(U, DT, 5, SEED, 3)
    errs["hmc_links"] = rel(dc.gather(c.links_download(), 18), ro["U"])
    errs["hmc_h_new"] = abs(r["h_new"] - ro["h_new"]) / abs(ro["h_new"])
    accept_equal = bool(r["accepted"] == ro["accepted"] and r["gauss_steps"] == ro["gauss_steps"])
    c.links_upload(dc.scatter(U, 18))
    c.sweep_heatbath(SEED, 11)
    errs["heatbath_links"] = rel(dc.gather(c.links_download(), 18), o.sweep_heatbath(U, SEED, 11))
    out = {"ranks": world, "grid": pg, "global_extent": gext, "transport": dc.transport,
           "max_rel_err": max(errs.values()), "errors": errs, "accept_equal": accept_equal,
           "gauss_steps": [int(r["gauss_steps"]), int(ro["gauss_steps"])],
           "tolerance": "deterministic paths 1e-12, heat bath 1e-9 (libm vs CUDA transcendentals)",
           "ok": bool(accept_equal and max(v for k, v in errs.items() if k != "heatbath_links") <= 1e-12
                      and errs["heatbath_links"] <= 1e-9)}
    dist.barrier()
    c.close()
    return out


def run_ours(args):
    # exactly ONE JSON line may reach stdout: NCCL/torch banners printed by native code go to stderr instead
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the product path has no CPU fallback"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L = args.extent
    D = 4
    scaling = "weak"
    if args.config == "d3":
        D, gext_cfg = 3, [40, 40, 40]
        assert world == 1, "--config d3 is a single-GPU case"
    elif args.config == "c4":
        gext_cfg, scaling = [48, 48, 48, 96], "strong"
    elif args.config == "c5":
        gext_cfg, scaling = [64, 64, 64, 64], "strong"
    else:
        gext_cfg = None
    parity = None
    numa_node = None
    if world > 1:
        from lattice_qcd_rs_b200.dist import bind_to_gpu_numa_node
        numa_node = bind_to_gpu_numa_node(local)  # node-local pinned buffers for the e2e copies of the N ranks
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        from lattice_qcd_rs_b200.dist import DistContext, proc_grid_for
        pg = proc_grid_for(D, world)
        parity = dist_parity(world, pg)
        gext = gext_cfg if gext_cfg is not None else [L * p for p in pg]
        dc = DistContext(D, gext, a=SPACING, beta=BETA, proc_grid=pg)
        ctx = dc.ctx
        stream = torch.cuda.current_stream(dev)
        par = ("x".join(str(p) for p in pg) + " ranks (z,t split), one-site halos: " +
               ("boundary slices stored into the neighbours' ghost layers over NVLink by our own kernels (CUDA IPC "
                "peer memory, release/acquire flags)" if dc.transport == "p2p" else "NCCL send/recv through callbacks"))
    else:
        from lattice_qcd_rs_b200 import Context
        gext = gext_cfg if gext_cfg is not None else [L] * 4
        ctx = Context(D, gext, a=SPACING, beta=BETA, device=local)
        stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
        dc, par, pg = None, "single GPU", [1] * D
    nl_local = ctx.nl
    nl_global = nl_local * world
    ns_local = nl_local // D

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(k):
            fn(i)
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---------------- synthetic hot start, resident in HBM (Philox random SU(3), decomposition independent); the
    # snapshot every timed step restarts from
    if args.flags:
        ctx.set_flags(args.flags)  # A/B switch for kernel variants (include/lqcd_b200.h LQ_FLAG_*); 0 = the defaults
    fp64_peak, copy_peak = ctx.measure_peaks()
    ctx.links_set_random(SEED, 0)
    ctx.efield_set_zero()
    ctx.snapshot()
    traj = {"n": 0, "acc": 0, "gauss": 0, "gauss_list": []}
    p2p0 = ctx.p2p_exchanges

    def trajectory(i):
        ctx.restore()
        r = ctx.hmc_trajectory(DT, MD_STEPS, SEED, TRAJ_COUNTER)
        ctx.reunitarize()
        traj["n"] += 1
        traj["acc"] += int(r["accepted"])
        traj["gauss"] += r["gauss_steps"]
        traj["gauss_list"].append(int(r["gauss_steps"]))

    for i in range(args.warmup):
        trajectory(i)
    ctx.profile_enable(True)
    sampler = ClockSampler(local)
    launches0 = ctx.kernel_launches
    gauss0 = traj["gauss"]
    traj["gauss_list"] = []
    p2p0 = ctx.p2p_exchanges
    if rank == 0:
        sampler.start()
    ms = timed(trajectory, args.steps)
    clocks = sampler.stop() if rank == 0 else {}
    launches = ctx.kernel_launches - launches0
    exchanges = (ctx.p2p_exchanges - p2p0) / max(args.steps, 1)
    gauss_steps = (traj["gauss"] - gauss0) / max(args.steps, 1)
    prof = {k: ctx.profile_get(k) for k in ("efield_link_step", "efield_step", "gauss_field", "gauss_step", "gauss_div",
                                            "plaquette", "efield_energy", "momenta", "reunitarize", "copy")}
    ctx.profile_enable(False)
    n_fused, ms_fused = prof["efield_link_step"]
    n_gf, ms_gf = prof["gauss_field"]
    n_gs, ms_gs = prof["gauss_step"]
    value = MD_STEPS * nl_global * args.steps / (ms * 1e-3)
    plaq = ctx.average_trace_plaquette().real / 3.0

    # ---------------- MD-only (no refresh / projection / H): lq_symplectic_n alone, for the breakdown
    ctx.restore()
    ctx.momenta_refresh(SEED, 77)
    ms_md = timed(lambda i: ctx.symplectic_n(DT, MD_STEPS), 1)
    md_only = MD_STEPS * nl_global / (ms_md * 1e-3)

    # ---------------- end to end through the C ABI with host buffers (pinned), copies inside the timed region
    ctx.restore()
    hostU = torch.empty((nl_local, 18), dtype=torch.float64, pin_memory=True)
    hostOut = torch.empty((nl_local, 18), dtype=torch.float64, pin_memory=True)
    hU, hOut = hostU.numpy(), hostOut.numpy()
    ctx.links_download(out=hU)
    e2e_state = {"n": 0, "gauss": 0}

    def e2e_step(i):
        ctx.links_upload(hU)
        e2e_state["gauss"] += ctx.hmc_trajectory(DT, MD_STEPS, SEED, TRAJ_COUNTER)["gauss_steps"]
        ctx.reunitarize()
        ctx.links_download(out=hOut)
        e2e_state["n"] += 1

    e2e_step(0)
    e2e_state["gauss"] = 0
    ms_e2e_serial = timed(e2e_step, args.steps)
    e2e_serial_value = MD_STEPS * nl_global * args.steps / (ms_e2e_serial * 1e-3)
    bytes_links = nl_local * 18 * 8

    # the same steps through the pipelined marshalling calls (lq_links_upload_begin / _commit, lq_links_download_begin,
    # lq_copies_wait): every step still uploads its input from and downloads its result to pinned host memory inside the
    # timed region, but the upload of step k + 1 and the download of step k travel while a trajectory computes (PCIe is
    # idle during compute and full duplex) -- the call sequence of a host that streams a batch of configurations
    def e2e_pipeline(i, n=None):
        n = args.steps if n is None else n
        ctx.links_upload_begin(hU)
        for k in range(n):
            ctx.links_upload_commit()
            if k + 1 < n:
                ctx.links_upload_begin(hU)
            e2e_state["gauss"] += ctx.hmc_trajectory(DT, MD_STEPS, SEED, TRAJ_COUNTER)["gauss_steps"]
            ctx.reunitarize()
            ctx.links_download_begin(hOut)
            e2e_state["n"] += 1
        ctx.copies_wait()

    e2e_pipeline(0, 1)  # untimed: allocates the staging buffers and copy streams
    e2e_state["gauss"] = 0
    ms_e2e = timed(e2e_pipeline, 1)
    e2e_value = MD_STEPS * nl_global * args.steps / (ms_e2e * 1e-3)
    hOut_check = float(np.abs(hOut).max())  # the last result is in host memory
    # the host <-> device copies alone (upload + download of the links, all ranks at once): what separates e2e from value

    def copy_step(i):
        ctx.links_upload(hU)
        ctx.links_download(out=hOut)

    ms_copy = timed(copy_step, 3) / 3

    # ---------------- local-update sweeps of config 3 (secondary numbers)
    sweeps = {}
    if world == 1 and D == 4:
        ctx.restore()
        for name, fn in (("heatbath", lambda i: ctx.sweep_heatbath(SEED, 9000 + i)),
                         ("overrelax", lambda i: ctx.sweep_overrelax(0)),
                         ("overrelax_su2_subgroups", lambda i: ctx.sweep_overrelax(2)),
                         ("metropolis", lambda i: ctx.sweep_metropolis(SEED, 9500 + i))):
            fn(0)
            t = timed(fn, 3)
            sweeps[name] = {"link_updates_per_s": nl_global * 3 / (t * 1e-3), "ms_per_sweep": t / 3,
                            "hbm_frac_algorithmic": (nl_global * 3 * 1296 / (t * 1e-3)) / 1e9 / peaks()[0]}
        # the same sweeps on a THERMALISED lattice (20 heat-bath sweeps from the hot start first): the state production
        # sweeps run on.  The Kennedy-Pendleton rejection loop runs at warp level, so its cost depends on the state: a
        # hot lattice (small staple determinants) rejects ~4 % of the candidates per lane, ~70 % of the warps repeat.
        ctx.restore()
        for i in range(20):
            ctx.sweep_heatbath(SEED, 9100 + i)
        therm_plaq = ctx.average_trace_plaquette().real / 3.0
        for name, fn in (("heatbath", lambda i: ctx.sweep_heatbath(SEED, 9200 + i)),
                         ("overrelax", lambda i: ctx.sweep_overrelax(0)),
                         ("metropolis", lambda i: ctx.sweep_metropolis(SEED, 9700 + i))):
            fn(0)
            t = timed(fn, 3)
            sweeps[name + "_thermalised"] = {"link_updates_per_s": nl_global * 3 / (t * 1e-3), "ms_per_sweep": t / 3,
                                             "hbm_frac_algorithmic": (nl_global * 3 * 1296 / (t * 1e-3)) / 1e9 / peaks()[0],
                                             "plaquette_over_3_before": therm_plaq}

    per_rank = None
    if world > 1:
        # per-rank averages of the two dominant kernels: the spread is the part of the multi-GPU loss that is GPU-to-GPU
        # clock variation under the power cap (every ghost refresh waits for the slowest rank)
        mine = torch.tensor([ms_fused / max(n_fused, 1), ms_gs / max(n_gs, 1)], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"md_kernel_avg_ms": [float(t[0]) for t in allr], "gauss_iteration_avg_ms": [float(t[1]) for t in allr],
                    "note": "each rank's own compute time (the waits for the neighbours sit in the barrier kernels between "
                            "launches); with --flags 512 / 2560 (halo synchronisation folded into the kernels) a kernel's "
                            "time includes its boundary blocks' wait"}
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    achieved = BYTES_FUSED * nl_local * n_fused / (ms_fused * 1e-3) / 1e9 if ms_fused > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("extent") == L and args.config == "metric":
            traffic = tj.get("dram_bytes_per_launch")
    cpu = None
    if world == 1 and args.config == "metric":  # reported on rank 0 at N = 1 only
        g_cpu = sample_gauss(SAMPLE_MD, gauss_steps)  # the mix of THIS run's trajectory
        cpu = cpu_sample(L, steps=1, warmup=0, gauss=g_cpu)[0]
        opt = cpu_sample(L, steps=1, warmup=0, literal=False, gauss=g_cpu)[0]
        cpu["optimised_cpu"] = {"value": opt["value"], "unit": opt["unit"], "sample": opt["sample"]}

    def _roof(kernel, bytes_per_launch, n, ms_tot):
        a = bytes_per_launch * n / (ms_tot * 1e-3) / 1e9 if ms_tot > 0 else 0.0
        return {"kernel": kernel, "launches": n, "avg_launch_ms": ms_tot / max(n, 1), "achieved": a, "frac": a / peak,
                "unit": "GB/s"}

    if D == 4 and not (args.flags & (4 | 8 | 32)):
        # projection loop on the transported field (default on D = 4): per site U 576 + E 256 + T 320 read, E' 256 +
        # T' 320 written by the iteration kernel; U 576 + E 256 read, T 320 written by the one-off initialisation
        gauss_rooflines = [_roof("lq_gtinit4_kernel (T = U^+ E U, once per projection, 1152 B/site)", 1152 * ns_local,
                                 n_gf, ms_gf),
                           _roof("lq_gausst4_kernel (project_to_gauss_step + EField::gauss in one pass, 1728 B/site)",
                                 1728 * ns_local, n_gs, ms_gs)]
    else:
        gauss_rooflines = [_roof("Gauss field (EField::gauss, 976 B/site)", 976 * ns_local, n_gf, ms_gf),
                           _roof("Gauss projection step (project_to_gauss_step, 308 B/link)", 308 * nl_local, n_gs, ms_gs)]
    steps = max(args.steps, 1)
    breakdown = {"fused_force_link_kernel": ms_fused / steps, "closing_force_kick": prof["efield_step"][1] / steps,
                 "gauss_field": ms_gf / steps, "gauss_project_step": ms_gs / steps,
                 "gauss_residual_reduce": prof["gauss_div"][1] / steps, "plaquette_reduce": prof["plaquette"][1] / steps,
                 "efield_energy_reduce": prof["efield_energy"][1] / steps, "momenta_refresh": prof["momenta"][1] / steps,
                 "reunitarize": prof["reunitarize"][1] / steps, "reject_path_copy": prof["copy"][1] / steps}
    breakdown["sum_of_kernels"] = sum(breakdown.values())
    breakdown["total"] = ms / steps
    breakdown["unaccounted_host_sync_and_restore"] = breakdown["total"] - breakdown["sum_of_kernels"]
    line = {
        "metric": "HMC link-updates/sec at 32^4 f64", "value": value, "unit": "link-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": workload_text(args.config, L),
            "config": args.config, "global_extent": gext, "local_extent": ctx.extent, "proc_grid": pg, "parallelism": par,
            "l2": "inputs (604 MB links + 268 MB E-field per GPU at 32^4) exceed the 126 MB L2; no flush needed",
            "md_steps_per_trajectory": MD_STEPS,
            "gauss_projection_steps_per_trajectory": gauss_steps, "gauss_steps_each": traj["gauss_list"],
            "gauss_steps_pinned_constant": GAUSS_STEPS_PINNED, "accept_rate": traj["acc"] / max(traj["n"], 1),
            "plaquette_over_3": plaq, "ghost_exchanges_per_step": exchanges, "rank0_numa_node": numa_node,
        },
        "md_only": {"value": md_only, "unit": "link-updates/s", "what": "lq_symplectic_n alone (no refresh/projection/H)"},
        "breakdown_ms_per_step": breakdown,
        "roofline": {"bound": "hbm", "kernel": "lq_md4_kernel<128,3,1> (fused force + E kick + link step" +
                     (" + halo push into the neighbours' ghost layers)" if world > 1 else ")"), "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_link": BYTES_FUSED,
                     # the same launches by SURVEY section 8(d)'s per-unit figure for the UNFUSED reference loops (one
                     # link-update = 2 x 272 B force+kick passes + 352 B link step = 896 B): what fusing the link step
                     # and merging the half-kicks saved shows up as measured traffic BELOW this figure
                     "survey_8d_unfused_accounting": {
                         "bytes_per_link_update": 896,
                         "achieved": 896 * nl_local * n_fused / (ms_fused * 1e-3) / 1e9 if ms_fused > 0 else 0.0,
                         "frac": (896 * nl_local * n_fused / (ms_fused * 1e-3) / 1e9 / peak) if ms_fused > 0 else 0.0},
                     "launches": n_fused,
                     "avg_launch_ms": ms_fused / max(n_fused, 1),
                     "fp64_tflops": FLOPS_FUSED * nl_local * n_fused / (ms_fused * 1e-3) / 1e12 if ms_fused > 0 else 0.0,
                     "fp64_peak_measured_in_run": fp64_peak, "copy_peak_measured_in_run_gbs": copy_peak,
                     "fp64_frac": (FLOPS_FUSED * nl_local * n_fused / (ms_fused * 1e-3) / 1e12 / fp64_peak)
                     if ms_fused > 0 and fp64_peak > 0 else None,
                     "fp64_note": "f64 FMA pipe is the tighter ceiling: ~3.1 kflop per 416 algorithmic bytes = 7.6 flop/B "
                                  "against a ridge of ~5.4 flop/B; DFMAs with three fresh register operands issue at 0.70 "
                                  "of the peak measured here (24.7 TFLOP/s, tools/kbench2.cu), the product code's own "
                                  "instruction stream at ~0.78"},
        "secondary_rooflines": gauss_rooflines,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "link-updates/s", "h2d_bytes_per_step": bytes_links,
                "d2h_bytes_per_step": bytes_links + 64, "ms_per_step": ms_e2e / args.steps,
                "how": "pipelined marshalling calls of the C ABI: the upload of step k+1 and the download of step k overlap "
                       "the trajectory of step k (two staging buffers, two copy streams); all copies inside the timed region",
                "serial_calls": {"value": e2e_serial_value, "ms_per_step": ms_e2e_serial / args.steps,
                                 "how": "lq_links_upload -> trajectory -> lq_links_download, each call synchronous"},
                "max_abs_of_last_result": hOut_check,
                "gauss_projection_steps_per_trajectory": e2e_state["gauss"] / max(args.steps, 1),
                "host_copies_alone": {"ms_per_step": ms_copy, "gb_per_s_per_gpu_each_way": 2 * bytes_links / (ms_copy * 1e-3) / 1e9 / 2,
                                      "what": "lq_links_upload + lq_links_download of the pinned host arrays on all ranks at "
                                              "once, transposition kernels included, max over ranks"}},
        "gpu_launches": launches, "flags": args.flags,
        "clocks": clocks,
        "sweeps": sweeps,
    }
    if parity is not None:
        line["parity"] = parity
    if per_rank is not None:
        line["per_rank"] = per_rank
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--extent", type=int, default=32)
    ap.add_argument("--config", default="metric", choices=["metric", "c4", "c5", "d3"],
                    help="metric: the BASELINE metric (32^4 per GPU, default); c4 / c5 / d3: BASELINE configs 4, 5, 5b")
    ap.add_argument("--flags", type=int, default=0, help="LQ_FLAG_* bits for A/B runs of kernel variants (default 0)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3  # timing hygiene: W >= 3
        run_ours(args)


if __name__ == "__main__":
    main()
