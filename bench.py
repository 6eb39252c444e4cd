#!/usr/bin/env python
"""bench.py -- HMC link-updates/sec at 32^4, f64, on N B200s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched by torch.distributed.run, one rank per GPU; weak scaling: 32^4 per GPU, split along t, and z at 8)

A "step" is one HMC trajectory of HybridMonteCarloDiagnostic::next_element (hybrid_monte_carlo.rs:465-471, 573-613) on
a hot-start 32^4 beta=6 lattice: momentum refresh -> Gauss projection -> H_old -> 100 symplectic-Euler MD steps
(dt = 0.01, the reference's own HMC bench shape, benches/bench.rs:166-189) -> H_new -> accept/reject, followed by
normalize_link_matrices.  One link-update = one link carried through one MD step, so a step is 100 * N_links
link-updates and ALL the per-trajectory overhead is inside the timed region.

value : device-resident (links stay in HBM between trajectories), CUDA events on the context stream, max over ranks.
e2e   : the same trajectory through the C ABI with HOST buffers: links uploaded from pinned host memory before and
        downloaded after every trajectory, inside the timed region.
--impl reference : the CPU restatement of the reference loops (oracle/, OpenMP over all host cores; the Rust crate
        cannot be built in this image) on a bounded sample of the same 32^4 workload (see CpuSample).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 0x457893F44AB067F0
BETA, SPACING, DT, MD_STEPS = 6.0, 1.0, 0.01, 100
# algorithmic bytes per link of the dominant kernel (fused force + E kick + link step), DESIGN.md section 4:
# read U 144 + read E 64 + write E 64 + write U' 144
BYTES_FUSED = 416
# its f64 flops per link: 12 staple matmuls + U*A (13 x 216) + trace/kick (~60) + E.to_matrix * U + Euler (~280)
FLOPS_FUSED = 13 * 216 + 60 + 280


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
# sample step is 1/25 of one -- SAMPLE_MD symplectic steps and SAMPLE_GAUSS Gauss-projection iterations (the 100 : 173
# mix of the GPU trajectory on this start configuration) on the full 32^4 hot lattice, literal reference loops.  What a
# sample leaves out (momentum refresh, 2 x H_total, accept, normalise) is < 1 % of a trajectory on either side.
SAMPLE_MD, SAMPLE_GAUSS = 4, 7


class CpuSample:
    def __init__(self, ext, md=SAMPLE_MD, gauss=SAMPLE_GAUSS):
        from oracle.oracle import Oracle
        self.o = o = Oracle(4, ext, a=SPACING, beta=BETA)
        o.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1; this arm uses every host core
        self.U = o.links_random(SEED)
        self.E = o.momenta_refresh(SEED, 1)
        self.ext, self.md, self.gauss = ext, md, gauss

    def step(self):
        o = self.o
        for _ in range(self.gauss):
            self.E = o.project_to_gauss_step(self.U, self.E)
        self.U, self.E = o.integrate(self.U, self.E, "symplectic", DT, n=self.md, literal=True)

    def describe(self):
        return (f"each step = {self.md} symplectic-Euler MD steps (dt={DT}) + {self.gauss} Gauss-projection iterations "
                f"on the full {self.ext}^4 beta={BETA} hot lattice (the GPU trajectory's 100 : 173 mix scaled down; "
                "momentum refresh, 2x H_total, accept and normalise left out: < 1 % of a trajectory), literal reference "
                "loops incl. the 28-matmul derivative_e, g++ -O3 -fopenmp on all host cores; C++ restatement of "
                "lattice-qcd-rs v0.2.1 (oracle/), not the Rust binary")


def cpu_sample(ext, steps=1, warmup=0, budget_s=240.0):
    """Oracle (C++/OpenMP restatement of the reference loops, literal mode) timed on the host cores.  The sample per step
    shrinks (4+7 -> 2+4 -> 1+2 MD steps + Gauss iterations) when steps x (time of one step) would exceed `budget_s`."""
    warmup = min(warmup, 1)  # no clock ramp or JIT on the CPU side: one step faults the pages in
    cs = CpuSample(ext)
    t0 = time.perf_counter()
    cs.step()  # calibration (counts as the warm-up step)
    t1 = time.perf_counter() - t0
    for md, gauss in ((2, 4), (1, 2)):
        if t1 * cs.md / SAMPLE_MD * steps > budget_s:
            cs.md, cs.gauss = md, gauss
    if warmup == 0 and steps == 1 and (cs.md, cs.gauss) == (SAMPLE_MD, SAMPLE_GAUSS):
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
    val = cpu["value"]
    line = {
        "impl": "reference", "metric": "HMC link-updates/sec at 32^4 f64", "value": val, "unit": "link-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.extent}^4 beta={BETA} HMC (100 MD steps/trajectory), hot start",
                   "sample": cpu["sample"]},
        "cpu_baseline": cpu,
        "e2e": {"value": val, "unit": "link-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        from lattice_qcd_rs_b200.dist import DistContext, proc_grid_for
        pg = proc_grid_for(4, world)
        gext = [L * p for p in pg]
        dc = DistContext(4, gext, a=SPACING, beta=BETA, proc_grid=pg)
        ctx = dc.ctx
        stream = torch.cuda.current_stream(dev)
        par = ("x".join(str(p) for p in pg) + " ranks (z,t split), one-site halos: " +
               ("boundary slices stored into the neighbours' ghost layers over NVLink by our own kernels (CUDA IPC "
                "peer memory, release/acquire flags)" if dc.transport == "p2p" else "NCCL send/recv through callbacks"))
    else:
        from lattice_qcd_rs_b200 import Context
        ctx = Context(4, L, a=SPACING, beta=BETA, device=local)
        stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
        gext, dc, par = [L] * 4, None, "single GPU"
    nl_local = ctx.nl
    nl_global = nl_local * world

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

    # ---------------- synthetic hot start, resident in HBM (Philox random SU(3), decomposition independent)
    if args.flags:
        ctx.set_flags(args.flags)  # A/B switch for kernel variants (include/lqcd_b200.h LQ_FLAG_*); 0 = the defaults
    ctx.links_set_random(SEED, 0)
    traj = {"n": 0, "acc": 0, "gauss": 0}

    def trajectory(i):
        r = ctx.hmc_trajectory(DT, MD_STEPS, SEED, 1 + traj["n"])
        ctx.reunitarize()
        traj["n"] += 1
        traj["acc"] += int(r["accepted"])
        traj["gauss"] += r["gauss_steps"]

    for i in range(args.warmup):
        trajectory(i)
    ctx.profile_enable(True)
    sampler = ClockSampler(local)
    launches0 = ctx.kernel_launches
    gauss0 = traj["gauss"]
    if rank == 0:
        sampler.start()
    ms = timed(trajectory, args.steps)
    clocks = sampler.stop() if rank == 0 else {}
    launches = ctx.kernel_launches - launches0
    gauss_steps = (traj["gauss"] - gauss0) / max(args.steps, 1)
    n_fused, ms_fused = ctx.profile_get("efield_link_step")
    n_gf, ms_gf = ctx.profile_get("gauss_field")
    n_gs, ms_gs = ctx.profile_get("gauss_step")
    n_pl, ms_pl = ctx.profile_get("plaquette")
    ctx.profile_enable(False)
    value = MD_STEPS * nl_global * args.steps / (ms * 1e-3)
    plaq = ctx.average_trace_plaquette().real / 3.0

    # ---------------- MD-only (no refresh / projection / H): lq_symplectic_n alone, for the breakdown
    ctx.momenta_refresh(SEED, 77)
    ms_md = timed(lambda i: ctx.symplectic_n(DT, MD_STEPS), 1)
    md_only = MD_STEPS * nl_global / (ms_md * 1e-3)
    ctx.links_set_random(SEED, 0)

    # ---------------- end to end through the C ABI with host buffers (pinned), copies inside the timed region
    hostU = torch.empty((nl_local, 18), dtype=torch.float64, pin_memory=True)
    hU = hostU.numpy()
    hU[:] = ctx.links_download()
    e2e_state = {"n": 0, "gauss": 0}

    def e2e_step(i):
        ctx.links_upload(hU)
        e2e_state["gauss"] += ctx.hmc_trajectory(DT, MD_STEPS, SEED, 5000 + e2e_state["n"])["gauss_steps"]
        ctx.reunitarize()
        ctx.links_download(out=hU)
        e2e_state["n"] += 1

    e2e_step(0)
    e2e_state["gauss"] = 0
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = MD_STEPS * nl_global * args.steps / (ms_e2e * 1e-3)
    bytes_links = nl_local * 18 * 8

    # ---------------- local-update sweeps of config 3 (secondary numbers)
    sweeps = {}
    if world == 1:
        for name, fn, cls in (("heatbath", lambda i: ctx.sweep_heatbath(SEED, 9000 + i), "heatbath"),
                              ("overrelax", lambda i: ctx.sweep_overrelax(0), "overrelax"),
                              ("overrelax_su2_subgroups", lambda i: ctx.sweep_overrelax(2), "overrelax"),
                              ("metropolis", lambda i: ctx.sweep_metropolis(SEED, 9500 + i), "metropolis")):
            fn(0)
            t = timed(fn, 3)
            sweeps[name] = {"link_updates_per_s": nl_global * 3 / (t * 1e-3), "ms_per_sweep": t / 3,
                            "hbm_frac_algorithmic": (nl_global * 3 * 1296 / (t * 1e-3)) / 1e9 / peaks()[0]}

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
        if tj.get("extent") == L:
            traffic = tj.get("dram_bytes_per_launch")
    cpu = cpu_sample(L, steps=1, warmup=0)[0] if world == 1 else None  # reported on rank 0 at N = 1 only
    ns_local = nl_local // 4

    def _roof(kernel, bytes_per_launch, n, ms_tot):
        a = bytes_per_launch * n / (ms_tot * 1e-3) / 1e9 if ms_tot > 0 else 0.0
        return {"kernel": kernel, "launches": n, "avg_launch_ms": ms_tot / max(n, 1), "achieved": a, "frac": a / peak,
                "unit": "GB/s"}

    if n_gs > 4 * max(n_gf, 1):  # one-pass iteration (D = 4 default): the Gauss field kernel runs once per projection
        gauss_rooflines = [_roof("lq_gauss4_kernel<128,3> (projection step + Gauss field of the projected E in one "
                                 "pass, 1376 B/site; FP64 co-limited: 32 matrix products/site)", 1376 * ns_local, n_gs, ms_gs),
                           _roof("KGaussField<4> (976 B/site)", 976 * ns_local, n_gf, ms_gf)]
    else:
        gauss_rooflines = [_roof("lq_gfield4_kernel (EField::gauss, 976 B/site)", 976 * ns_local, n_gf, ms_gf),
                           _roof("lq_gstep4_kernel (project_to_gauss_step, 308 B/link)", 308 * nl_local, n_gs, ms_gs)]
    line = {
        "metric": "HMC link-updates/sec at 32^4 f64", "value": value, "unit": "link-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": f"{L}^4 per GPU, beta={BETA}, a={SPACING}: full HMC trajectory (refresh sigma=0.5/beta, Gauss "
                        f"projection, 2x H_total, {MD_STEPS} symplectic-Euler steps dt={DT}, accept/reject) + "
                        "normalize_link_matrices; hot start (Philox random SU(3))",
            "global_extent": gext, "parallelism": par, "l2": "inputs (604 MB links + 268 MB E-field per GPU) exceed the "
            "126 MB L2; no flush needed", "md_steps_per_trajectory": MD_STEPS,
            "gauss_projection_steps_per_trajectory": gauss_steps, "accept_rate": traj["acc"] / max(traj["n"], 1),
            "plaquette_over_3": plaq,
        },
        "md_only": {"value": md_only, "unit": "link-updates/s", "what": "lq_symplectic_n alone (no refresh/projection/H)"},
        "breakdown_ms_per_step": {"fused_force_link_kernel": ms_fused / args.steps, "gauss_field": ms_gf / args.steps,
                                  "gauss_project_step": ms_gs / args.steps, "plaquette_reduce": ms_pl / args.steps,
                                  "total": ms / args.steps},
        "roofline": {"bound": "hbm", "kernel": "lq_md4_kernel<128,3,1,2> (fused force + E kick + link step" +
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
                     "fp64_note": "f64 FMA pipe is the tighter ceiling: ~3.1 kflop per 416 algorithmic bytes = 7.6 flop/B "
                                  "against a ridge of 5.4 flop/B (35.3 TF measured / 6.53 TB/s); frac vs HBM cannot exceed "
                                  "0.71"},
        "secondary_rooflines": gauss_rooflines,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "link-updates/s", "h2d_bytes_per_step": bytes_links,
                "d2h_bytes_per_step": bytes_links + 64, "ms_per_step": ms_e2e / args.steps,
                "gauss_projection_steps_per_trajectory": e2e_state["gauss"] / max(args.steps, 1)},
        "gpu_launches": launches, "flags": args.flags,
        "clocks": clocks,
        "sweeps": sweeps,
    }
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
