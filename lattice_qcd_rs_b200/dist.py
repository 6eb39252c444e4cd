"""Domain decomposition plumbing: one process per GPU, torch.distributed (NCCL over NVLink on GPUs).

The lattice is split along the slowest direction t = x_{D-1} (and z = x_{D-2} when the process grid asks for
it); each rank owns a block plus one-site-deep ghost layers.  The CUDA library sequences
every kernel itself and asks this module for exactly two services through `lq_set_comm`:
  * halo_exchange(which): pack the two boundary slices of every split direction with the library's pack kernel,
    exchange them with the +-1 neighbours (grouped isend/irecv), unpack into the ghost layers;
  * allreduce_sum(vals): global sums of a few f64 (plaquette, Hamiltonians, Gauss residual, accept statistics).
Nothing here computes on the lattice; torch is device memory + transport only.
"""
import itertools
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

from ._capi import Context


def proc_grid_for(D, world):
    """t-slabs up to 4 ranks, 4(t) x 2(z) at 8."""
    pg = [1] * D
    if world <= 4 or D < 3:
        pg[D - 1] = world
    else:
        pg[D - 1] = world // 2
        pg[D - 2] = 2
    return pg


def bind_to_gpu_numa_node(gpu_index):
    """Pin this process (one rank per GPU) to the CPU cores of the NUMA node its GPU hangs off, so that the pinned host
    buffers it allocates afterwards are node-local (first touch) and the eight ranks' host<->device copies do not all
    cross the socket interconnect.  Returns the NUMA node, or None when the topology cannot be read (no change made)."""
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(gpu_index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().splitlines()[0].strip().lower()
        # nvidia-smi prints an 8-digit domain (00000000:1B:00.0); sysfs uses 4 digits
        dom, rest = out.split(":", 1)
        dev = f"/sys/bus/pci/devices/{dom[-4:]}:{rest}"
        node = int(open(dev + "/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:  # noqa: BLE001 -- an optimisation only
        return None


class DistContext:
    def __init__(self, D, global_extent, a=1.0, beta=1.0, CA=3.0, proc_grid=None, lib=None, device=None, group=None):
        assert dist.is_initialized(), "torch.distributed must be initialised (one process per GPU)"
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if np.isscalar(global_extent):
            global_extent = [int(global_extent)] * D
        self.D = D
        self.global_extent = list(global_extent)
        self.proc_grid = list(proc_grid) if proc_grid is not None else proc_grid_for(D, self.world)
        assert int(np.prod(self.proc_grid)) == self.world
        assert all(p == 1 for p in self.proc_grid[:max(D - 2, 0)]), "only the two slowest directions may be split"
        # rank -> coordinates: the fastest-varying split direction first
        self.coord = [0] * D
        r = self.rank
        for d in range(D):
            self.coord[d] = r % self.proc_grid[d]
            r //= self.proc_grid[d]
        self.on_cuda = lib is None
        if self.on_cuda:
            self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
            dev_index = self.device.index
        else:  # host-emulation library (CPU CI over gloo)
            self.device = torch.device("cpu")
            dev_index = 0
        self.ctx = Context(D, self.global_extent, a=a, beta=beta, CA=CA, device=dev_index, lib=lib,
                           proc_grid=self.proc_grid, rank_coord=self.coord)
        if self.on_cuda:
            # all library work is ordered on torch's current stream, which torch's NCCL ops synchronise with
            self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        self._bufs = {}
        self._views = {}
        self._scratch = torch.zeros(16, dtype=torch.float64, device=self.device)
        self.halo_exchanges = 0
        self.halo_bytes_sent = 0
        self.transport = "none"
        if self.world > 1:
            self.ctx.set_comm(self._halo_exchange, self._allreduce_sum, self._allreduce_sum_device)
            self.transport = "nccl-callbacks"
            if self.on_cuda and os.environ.get("LQ_HALO_TRANSPORT", "p2p") == "p2p":
                self._setup_p2p()

    # ------------------------------------------------------------------ peer-to-peer transport (CUDA IPC over NVLink)
    def neighbour_offsets(self):
        """All offsets in {-1,0,1} over the split directions except 0, in an order every rank agrees on."""
        dec = [d for d in range(self.D) if self.proc_grid[d] > 1]
        out = []
        for combo in itertools.product((-1, 0, 1), repeat=len(dec)):
            if any(combo):
                o = [0] * self.D
                for d, v in zip(dec, combo):
                    o[d] = v
                out.append(o)
        return out

    def _setup_p2p(self):
        """Exchange IPC handles of the field buffers once; afterwards ghost refreshes are the library's own kernels
        storing into the neighbours' memory.  Falls back to the NCCL callbacks when IPC/peer access is unavailable
        (every rank takes the same decision)."""
        ok, mine = 1, b""
        try:
            mine = self.ctx.p2p_export()
        except Exception as e:  # noqa: BLE001
            print(f"[lattice_qcd_rs_b200] rank {self.rank}: p2p export failed ({e}); using NCCL halos", file=sys.stderr)
            ok = 0
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (ok, mine), group=self.group)
        if not all(g[0] for g in gathered):
            return
        offs = self.neighbour_offsets()
        ranks = [self._rank_of([c + o for c, o in zip(self.coord, off)]) for off in offs]
        peers = sorted(set(ranks))
        try:
            self.ctx.p2p_attach([gathered[r][1] for r in peers], offs, [peers.index(r) for r in ranks])
            ok = 1
        except Exception as e:  # noqa: BLE001
            print(f"[lattice_qcd_rs_b200] rank {self.rank}: p2p attach failed ({e}); using NCCL halos", file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            self.transport = "p2p"
        else:
            raise RuntimeError("peer-to-peer attach succeeded on some ranks only; set LQ_HALO_TRANSPORT=nccl")

    # ------------------------------------------------------------------ rank geometry
    def _rank_of(self, coord):
        r, mul = 0, 1
        for d in range(self.D):
            r += (coord[d] % self.proc_grid[d]) * mul
            mul *= self.proc_grid[d]
        return r

    def _neighbour(self, d, step):
        c = list(self.coord)
        c[d] += step
        return self._rank_of(c)

    # ------------------------------------------------------------------ callbacks
    def _halo_exchange(self, handle, which):
        # `handle` is the lq_ctx* the library wants refreshed: this rank's context or a clone of it (lq_ctx_clone copies
        # the callbacks, not the peer-to-peer mappings, so a clone's ghost layers always travel through here)
        ctx = self.ctx.borrowed(handle)
        for d in range(self.D):
            if not ctx.is_decomposed(d):
                continue
            nbytes = ctx.halo_bytes(which, d)
            key = (which, d)
            if key not in self._bufs:
                n = nbytes // 8
                self._bufs[key] = [torch.empty(n, dtype=torch.float64, device=self.device) for _ in range(4)]
            send_lo, send_hi, recv_lo, recv_hi = self._bufs[key]
            ctx.halo_pack(which, d, 0, send_lo.data_ptr(), nbytes)
            ctx.halo_pack(which, d, 1, send_hi.data_ptr(), nbytes)
            lo, hi = self._neighbour(d, -1), self._neighbour(d, +1)
            # my low face becomes the HIGH ghost of my low neighbour, and vice versa.  Posting order
            # [send_lo, send_hi] / [recv_hi, recv_lo] keeps the pairs matched when lo == hi (2 ranks).
            ops = [dist.P2POp(dist.isend, send_lo, lo, self.group, tag=0),
                   dist.P2POp(dist.isend, send_hi, hi, self.group, tag=1),
                   dist.P2POp(dist.irecv, recv_hi, hi, self.group, tag=0),
                   dist.P2POp(dist.irecv, recv_lo, lo, self.group, tag=1)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            ctx.halo_unpack(which, d, 1, recv_hi.data_ptr(), nbytes)
            ctx.halo_unpack(which, d, 0, recv_lo.data_ptr(), nbytes)
            self.halo_bytes_sent += 2 * nbytes
        self.halo_exchanges += 1
        return 0

    def _allreduce_sum(self, vals):
        n = vals.shape[0]
        t = self._scratch[:n]
        t.copy_(torch.from_numpy(vals))
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        vals[:] = t.cpu().numpy()
        return 0

    def _allreduce_sum_device(self, ptr, n):
        """In-place global sum of the library's own result buffer: a tensor view of that memory (device memory through
        __cuda_array_interface__, host memory for the CPU CI), one all-reduce enqueued on the library's stream."""
        key = (ptr, n)
        t = self._views.get(key)
        if t is None:
            if self.on_cuda:
                class _Raw:
                    __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}
                t = torch.as_tensor(_Raw(), device=self.device)
            else:
                import ctypes
                t = torch.from_numpy(np.ctypeslib.as_array((ctypes.c_double * n).from_address(ptr)))
            self._views[key] = t
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return 0

    # ------------------------------------------------------------------ scatter / gather of reference-order arrays
    def local_slices(self):
        sl = []
        for d in range(self.D):
            n = self.global_extent[d] // self.proc_grid[d]
            sl.append(slice(self.coord[d] * n, (self.coord[d] + 1) * n))
        return sl

    def scatter(self, global_array, per_site):
        """global (Ns_global * per_site...) array in reference order -> this rank's block in reference order."""
        g = np.asarray(global_array).reshape(*self.global_extent[::-1], -1)
        sl = self.local_slices()[::-1]
        return np.ascontiguousarray(g[tuple(sl)]).reshape(-1, per_site)

    def gather(self, local_array, per_site):
        """inverse of scatter; every rank gets the global array."""
        loc = torch.from_numpy(np.ascontiguousarray(local_array, dtype=np.float64)).to(self.device)
        parts = [torch.empty_like(loc) for _ in range(self.world)]
        dist.all_gather(parts, loc, group=self.group)
        out = np.empty(self.global_extent[::-1] + [local_array.size // self.ctx.ns], dtype=np.float64)
        for r, p in enumerate(parts):
            c, rr = [0] * self.D, r
            for d in range(self.D):
                c[d] = rr % self.proc_grid[d]
                rr //= self.proc_grid[d]
            sl = []
            for d in range(self.D):
                n = self.global_extent[d] // self.proc_grid[d]
                sl.append(slice(c[d] * n, (c[d] + 1) * n))
            block = p.cpu().numpy().reshape(*[s.stop - s.start for s in sl[::-1]], -1)
            out[tuple(sl[::-1])] = block
        return out.reshape(-1, per_site)
