"""x-slab domain decomposition of one MPMSimulation across the GPUs of a node.

Nothing in the reference corresponds to this (single address space, src/nclr.h:100-101).  One process per
GPU; rank r owns the particles whose stencil base `base.x = (int)(x*inv_dx - 0.5)` (src/nclr.h:115) lies in
`[bounds[r], bounds[r+1])`.  x is the slowest grid index (src/nclr.h:141), so a slab is a contiguous range of
the dense grid.  Per step (advance(), src/nclr.h:80-84, cut where neighbours must talk):

    engine.p2g()                    re-bin/sort, clear planes [x0, x1+2), P2G of the slab's particles
    exchange A                      both neighbours swap their partial sums of the 2 shared node planes and add
    engine.grid_g2p(...)            grid_op on [x0, x1+2) (shared planes redundantly, bit-identical on both
                                    sides), G2P, pack particles whose new base.x left the slab
    exchange B                      all_gather of the per-rank migrant counts, then the records left/right
    engine.unpack(...)              append received records

The transport is torch.distributed point-to-point (NCCL on GPUs; gloo in the CPU tests, where an oracle-backed
engine stands in for the device).  No collective touches the grid or particle data path except the 2×world
int32 count table.  Boundaries are chosen by particle count and can be re-balanced every k steps.
"""
from __future__ import annotations

import ctypes as ct

import numpy as np

from . import sim as _sim


# ------------------------------------------------------------------------------------------------
# host logic: ownership and count-balanced boundaries
# ------------------------------------------------------------------------------------------------
def base_x(x: np.ndarray, res: int) -> np.ndarray:
    """Stencil base along x exactly as the solver computes it (fp32 multiply, subtract, truncating cast)."""
    dx = np.float32(1.0 / res)
    inv_dx = np.float32(1.0) / dx
    g = x[:, 0].astype(np.float32) * inv_dx
    return np.trunc(g - np.float32(0.5)).astype(np.int32)


def balanced_bounds(hist: np.ndarray, world: int, min_width: int = 4) -> list[int]:
    """Slab boundaries b[0]=0 < ... < b[world]=len(hist) with ~equal particle counts per slab.

    hist[i] = number of particles with base.x == i.  Every slab is at least `min_width` planes wide so that the
    two ghost planes of a slab lie inside its right neighbour."""
    n1 = len(hist)
    if world * min_width > n1:
        raise ValueError(f"{world} slabs of >= {min_width} planes do not fit {n1} node planes")
    cum = np.concatenate([[0], np.cumsum(hist, dtype=np.int64)])
    total = int(cum[-1])
    b = [0]
    for k in range(1, world):
        target = total * k / world
        cut = int(np.searchsorted(cum, target, side="left"))  # first plane index with cum >= target
        cut = max(cut, b[-1] + min_width)
        cut = min(cut, n1 - (world - k) * min_width)
        b.append(cut)
    b.append(n1)
    return b


def load_imbalance(hist: np.ndarray, bounds: list[int]) -> float:
    """Largest relative deviation of a slab's particle count from the mean, for boundaries `bounds` over the per-plane
    histogram `hist` (0 = perfectly balanced)."""
    cum = np.concatenate([[0], np.cumsum(hist, dtype=np.int64)])
    loads = np.diff(cum[np.asarray(bounds)])
    mean = max(1.0, float(cum[-1]) / (len(bounds) - 1))
    return float(np.abs(loads - mean).max()) / mean


def limited_shift(old: list[int], new: list[int], max_shift: int, min_width: int = 4) -> list[int]:
    """Move every interior boundary towards `new` by at most `max_shift` planes, keeping min widths."""
    out = [old[0]]
    for k in range(1, len(old) - 1):
        c = int(np.clip(new[k], old[k] - max_shift, old[k] + max_shift))
        c = max(c, out[-1] + min_width)
        out.append(c)
    out.append(old[-1])
    for k in range(len(out) - 2, 0, -1):  # right-to-left pass for the upper constraint
        out[k] = min(out[k], out[k + 1] - min_width)
    return out


# ------------------------------------------------------------------------------------------------
# device engine: one slab on one GPU through the C-ABI
# ------------------------------------------------------------------------------------------------
class _DevArray:
    """Zero-copy view of library-owned device memory for torch.as_tensor."""

    def __init__(self, ptr: int, nfloats: int):
        self.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (ptr, False), "version": 3}


class GpuSlabEngine:
    """One x-slab on one GPU (libnmpm.so, nmpm_slab_* of include/nmpm.h)."""

    def __init__(self, x, ids, model, res, dt, E, nu, gravity, slab, capacity, device, sort_every=4, **state):
        import torch
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.sim = _sim.MPMSimulation(x, model, res, dt, E, nu, gravity, device=device, slab=slab, capacity=capacity,
                                      ids=ids, sort_every=sort_every, **state)
        self.sim.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        self._L, self._h = self.sim._L, self.sim._h
        self.dim, self.res = self.sim.dim, res
        self.rec_words = int(self._L.nmpm_migrate_record_bytes(self._h)) // 4
        self.plane_words = int(self._L.nmpm_grid_plane_bytes(self._h)) // 4

    def new_buffer(self, n, dtype="float32"):
        return self.torch.zeros(int(n), dtype=getattr(self.torch, dtype), device=self.device)

    def p2g(self):
        self.sim._check(self._L.nmpm_slab_p2g(self._h), "nmpm_slab_p2g")

    def plane_view(self, x_plane, planes):
        ptr = self._L.nmpm_grid_plane_ptr(self._h, int(x_plane))
        return self.torch.as_tensor(_DevArray(int(ptr), planes * self.plane_words), device=self.device)

    def add_planes(self, x_plane, planes, buf):
        self.sim._check(self._L.nmpm_grid_add_planes(self._h, int(x_plane), int(planes), ct.c_void_p(buf.data_ptr())),
                        "nmpm_grid_add_planes")

    def grid_g2p(self, send_left, send_right, cap_records, counts):
        self.sim._check(self._L.nmpm_slab_grid_g2p(self._h, ct.c_void_p(send_left.data_ptr()),
                                                   ct.c_void_p(send_right.data_ptr()), int(cap_records),
                                                   ct.c_void_p(counts.data_ptr())), "nmpm_slab_grid_g2p")

    def unpack(self, recv_left, n_left, recv_right, n_right, n_sent):
        self.sim._check(self._L.nmpm_slab_unpack(self._h, ct.c_void_p(recv_left.data_ptr()), int(n_left),
                                                 ct.c_void_p(recv_right.data_ptr()), int(n_right), int(n_sent)),
                        "nmpm_slab_unpack")

    def set_range(self, x0, x1):
        self.sim._check(self._L.nmpm_slab_set_range(self._h, int(x0), int(x1)), "nmpm_slab_set_range")

    def histogram(self, hist):
        self.sim._check(self._L.nmpm_slab_histogram(self._h, ct.c_void_p(hist.data_ptr())), "nmpm_slab_histogram")

    # -- native step: the same protocol inside libnmpm with NCCL calls on the sim's stream -------------
    @staticmethod
    def _libnccl_path() -> bytes:
        try:
            import nvidia.nccl
            from pathlib import Path
            p = Path(nvidia.nccl.__path__[0]) / "lib" / "libnccl.so.2"
            return str(p).encode() if p.exists() else b""
        except Exception:
            return b""

    def native_init(self, dist, group, rank, world, bounds, cap_records, n_global=None):
        """Create the library's NCCL communicator: rank 0 makes the unique id, torch.distributed carries it."""
        torch = self.torch
        buf = (ct.c_ubyte * 128)()
        path = self._libnccl_path()
        if rank == 0:
            self.sim._check(self._L.nmpm_nccl_unique_id(buf, path), "nmpm_nccl_unique_id")
        t = torch.tensor(list(buf), dtype=torch.uint8, device=self.device)
        dist.broadcast(t, src=0, group=group)
        raw = bytes(t.cpu().tolist())
        b = np.ascontiguousarray(bounds, np.int32)
        self.sim._check(self._L.nmpm_slab_comm_init(self._h, raw, rank, world, b.ctypes.data_as(_sim._i32p),
                                                    int(cap_records), path), "nmpm_slab_comm_init")
        if n_global is not None:
            self.sim._check(self._L.nmpm_slab_set_global_count(self._h, int(n_global)), "nmpm_slab_set_global_count")

    def native_step(self, nsteps):
        self.sim._check(self._L.nmpm_slab_step(self._h, int(nsteps)), "nmpm_slab_step")

    def native_set_bounds(self, bounds):
        b = np.ascontiguousarray(bounds, np.int32)
        self.sim._check(self._L.nmpm_slab_set_bounds(self._h, b.ctypes.data_as(_sim._i32p)), "nmpm_slab_set_bounds")

    def native_migrated(self):
        return int(self._L.nmpm_slab_migrated(self._h))

    def num_particles(self):
        """Live particles on this slab (device-driven native step: synchronises and raises on a latched step error)."""
        np_, ns_ = ct.c_longlong(), ct.c_longlong()
        self.sim._check(self._L.nmpm_slab_counts(self._h, ct.byref(np_), ct.byref(ns_)), "nmpm_slab_counts")
        return int(np_.value)

    def download_slots(self, out=None, compact=True):
        return self.sim.particles_slots(out, compact)

    def pinned_slot_buffers(self, slots):
        """Pinned host arrays for download_slots(out=...): the D2H copies then run at PCIe speed."""
        t, d = self.torch, self.dim
        shapes = dict(x=(slots, d), v=(slots, d), F=(slots, d, d), C=(slots, d, d), Jp=(slots,))
        self._pinned = {k: t.empty(s, dtype=t.float32).pin_memory() for k, s in shapes.items()}
        self._pinned["ids"] = t.empty((slots,), dtype=t.int32).pin_memory()
        out = {k: v.numpy() for k, v in self._pinned.items()}
        out["ids"] = out["ids"].view(np.uint32)
        return out

    def grid(self):
        return self.sim.grid()

    def synchronize(self):
        self.sim.synchronize()

    def launch_count(self):
        return self.sim.launch_count()


# ------------------------------------------------------------------------------------------------
# the driver (backend-agnostic: NCCL + GpuSlabEngine in production, gloo + an oracle engine in tests)
# ------------------------------------------------------------------------------------------------
class SlabSimulation:
    """MPMSimulation<dim> partitioned into world_size x-slabs; same constructor arguments as the reference class
    (src/nclr.h:74-78) plus the process group.  Every rank passes the SAME global particle arrays."""

    MIN_WIDTH = 4
    REBALANCE_TOL = 0.03  # relative load deviation below which the slab boundaries stay where they are

    def __init__(self, particles, model, res=64, dt=1e-4, E=1e4, nu=0.2, gravity=-100.0, *, group=None, device=0,
                 engine_factory=None, slack=0.5, rebalance_every=0, max_shift=2, bounds=None, native=None, **state):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        x = np.ascontiguousarray(particles, np.float32)
        self.n_global, self.dim = x.shape
        self.res, self.n1 = int(res), int(res) + 1
        self.rebalance_every, self.max_shift = int(rebalance_every), int(max_shift)
        bx = base_x(x, res)
        hist = np.bincount(np.clip(bx, 0, res), minlength=self.n1)
        self.bounds = list(bounds) if bounds is not None else balanced_bounds(hist, self.world, self.MIN_WIDTH)
        assert len(self.bounds) == self.world + 1 and self.bounds[0] == 0 and self.bounds[-1] == self.n1
        x0, x1 = self._own_range(self.bounds)
        mine = np.nonzero((bx >= x0) & (bx < x1))[0]
        n_local = len(mine)
        # room for migrants and for drift of the balance between re-balancing steps
        self.capacity = int(n_local * (1.0 + slack)) + 65536
        # migrant message capacity: the SAME on every rank (the native step sizes its fixed-capacity messages from it)
        self.cap_records = int(self.n_global / self.world * 0.25) + 65536
        sub = {k: (None if v is None else np.ascontiguousarray(v)[mine]) for k, v in state.items()}
        factory = engine_factory or GpuSlabEngine
        self.engine = factory(x[mine], mine.astype(np.uint32), int(model), int(res), dt, E, nu, gravity, (x0, x1),
                              self.capacity, device, **sub)
        e = self.engine
        self.W = e.rec_words
        self.hist = e.new_buffer(self.n1, "int32") if self.rebalance_every else None
        self.grid_bounds = list(self.bounds)   # boundaries the particles currently obey (exchange A)
        self.steps = 0
        self.migrated = 0
        # native = the protocol below runs inside libnmpm (NCCL send/recv issued from C++); default whenever the
        # engine offers it and the group is NCCL.  The Python protocol stays as the reference implementation
        # (gloo CPU tests, and `native=False` for A/B runs).
        if native is None:
            native = hasattr(e, "native_init") and dist.get_backend(group) == "nccl"
        self.native = bool(native)
        if self.native:
            e.native_init(dist, group, self.rank, self.world, self.bounds, self.cap_records, self.n_global)
            return
        self.buf_from_left = e.new_buffer(2 * e.plane_words)
        self.buf_from_right = e.new_buffer(2 * e.plane_words)
        self.send_left = e.new_buffer(self.cap_records * self.W)
        self.send_right = e.new_buffer(self.cap_records * self.W)
        self.recv_left = e.new_buffer(self.cap_records * self.W)
        self.recv_right = e.new_buffer(self.cap_records * self.W)
        self.counts = e.new_buffer(4, "int32")
        self.table = e.new_buffer(4 * self.world, "int32")

    # rank r owns base.x in [b[r], b[r+1]); the outermost slabs also keep whatever lies beyond the grid
    def _own_range(self, b):
        x0 = b[self.rank] if self.rank > 0 else 0
        x1 = b[self.rank + 1] if self.rank < self.world - 1 else max(self.n1, b[-1]) + (1 << 20)
        return x0, x1

    @property
    def left(self):
        return self.rank - 1 if self.rank > 0 else None

    @property
    def right(self):
        return self.rank + 1 if self.rank < self.world - 1 else None

    def _p2p(self, ops):
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()

    def _exchange_planes(self):
        d, e, g = self.dist, self.engine, self.group
        b = self.grid_bounds
        ops = []
        if self.right is not None:
            xr = b[self.rank + 1]
            ops.append(d.P2POp(d.isend, e.plane_view(xr, 2), self.right, g))
            ops.append(d.P2POp(d.irecv, self.buf_from_right, self.right, g))
        if self.left is not None:
            xl = b[self.rank]
            ops.append(d.P2POp(d.isend, e.plane_view(xl, 2), self.left, g))
            ops.append(d.P2POp(d.irecv, self.buf_from_left, self.left, g))
        self._p2p(ops)
        if self.right is not None:
            e.add_planes(b[self.rank + 1], 2, self.buf_from_right)
        if self.left is not None:
            e.add_planes(b[self.rank], 2, self.buf_from_left)

    def _exchange_migrants(self):
        d, e, g, W = self.dist, self.engine, self.group, self.W
        d.all_gather_into_tensor(self.table, self.counts, group=g)
        t = self.table.cpu().numpy().reshape(self.world, 4)  # the one host sync of the step
        if t[:, 3].any():
            raise RuntimeError(f"slab migration buffer overflow on rank(s) {np.nonzero(t[:, 3])[0].tolist()} "
                               f"(cap {self.cap_records} records)")
        nl, nr = int(t[self.rank, 0]), int(t[self.rank, 1])
        from_left = int(t[self.rank - 1, 1]) if self.left is not None else 0
        from_right = int(t[self.rank + 1, 0]) if self.right is not None else 0
        if (self.left is None and nl) or (self.right is None and nr):
            raise RuntimeError("a particle left the outermost slab")
        if max(from_left, from_right) > self.cap_records:
            raise RuntimeError("slab migration receive buffer too small")
        ops = []
        if nl:
            ops.append(d.P2POp(d.isend, self.send_left[:nl * W], self.left, g))
        if nr:
            ops.append(d.P2POp(d.isend, self.send_right[:nr * W], self.right, g))
        if from_left:
            ops.append(d.P2POp(d.irecv, self.recv_left[:from_left * W], self.left, g))
        if from_right:
            ops.append(d.P2POp(d.irecv, self.recv_right[:from_right * W], self.right, g))
        self._p2p(ops)
        e.unpack(self.recv_left, from_left, self.recv_right, from_right, nl + nr)
        self.migrated += nl + nr

    def _rebalance(self):
        d, e = self.dist, self.engine
        self.hist.zero_()
        e.histogram(self.hist)
        d.all_reduce(self.hist, group=self.group)
        hist = self.hist.cpu().numpy()
        # hysteresis: a boundary move makes up to max_shift planes of particles migrate in one step; leave slabs that
        # are within REBALANCE_TOL of the mean load alone
        if load_imbalance(hist, self.bounds) <= self.REBALANCE_TOL:
            return
        target = balanced_bounds(hist, self.world, self.MIN_WIDTH)
        new = limited_shift(self.bounds, target, self.max_shift, self.MIN_WIDTH)
        if new != self.bounds:
            self.bounds = new
            if self.native:
                e.native_set_bounds(new)
            else:
                e.set_range(*self._own_range(new))

    def step(self):
        if self.native:
            return self.advance(1)
        e = self.engine
        e.p2g()
        self._exchange_planes()
        e.grid_g2p(self.send_left, self.send_right, self.cap_records, self.counts)
        self.grid_bounds = list(self.bounds)  # after this G2P every particle obeys the current boundaries
        self.steps += 1
        if self.rebalance_every and self.steps % self.rebalance_every == 0:
            self._rebalance()                 # takes effect in the NEXT step's G2P
        self._exchange_migrants()

    def advance(self, nsteps=1):
        nsteps = int(nsteps)
        if not self.native:
            for _ in range(nsteps):
                self.step()
            return
        while nsteps > 0:  # native: whole runs of steps per call, cut only where a re-balance is due
            k = nsteps
            if self.rebalance_every:
                k = min(k, self.rebalance_every - self.steps % self.rebalance_every)
            self.engine.native_step(k)
            self.steps += k
            nsteps -= k
            self.migrated = self.engine.native_migrated()
            if self.rebalance_every and self.steps % self.rebalance_every == 0:
                self._rebalance()

    def synchronize(self):
        self.engine.synchronize()

    def num_local(self):
        return self.engine.num_particles()

    def particles(self, dst=0):
        """particles() of the global simulation in INPUT order, assembled on rank `dst` (None elsewhere)."""
        local = self.engine.download_slots()
        gathered = [None] * self.world if self.rank == dst else None
        self.dist.gather_object(local, gathered, dst=dst, group=self.group)
        if self.rank != dst:
            return None
        n, d = self.n_global, self.dim
        out = dict(x=np.empty((n, d), np.float32), v=np.empty((n, d), np.float32), F=np.empty((n, d, d), np.float32),
                   C=np.empty((n, d, d), np.float32), Jp=np.empty((n,), np.float32))
        seen = np.zeros(n, np.int32)
        for part in gathered:
            ids = part["ids"].astype(np.int64)
            seen[ids] += 1
            for k in out:
                out[k][ids] = part[k]
        if not (seen == 1).all():
            raise RuntimeError(f"slab bookkeeping lost or duplicated particles: {(seen != 1).sum()} ids off")
        return out


# ------------------------------------------------------------------------------------------------
# bench.py --gpus N
# ------------------------------------------------------------------------------------------------
def verify_slabs(x, model, res, rank, world, local, steps=3):
    """Correctness of the partitioned step, checked in the bench run itself: a sub-block of the scene (<= 128^3 particles,
    same spacing / material / grid) is advanced `steps` steps by all `world` slabs and by a single-GPU twin on rank 0;
    the gathered slab state must hold every particle exactly once and agree with the twin within the tolerances of
    tests/test_parity_gpu.py::test_full_size_properties_config4 (summation order differs across the slab cut).
    3 steps: 3D snow decorrelates at step 4 (Q1)."""
    import torch.distributed as dist
    from . import sim as _s
    dim = x.shape[1]
    n_sub = min(len(x), 128 ** 3)
    # a contiguous x-major prefix of the block keeps its spacing; trim to whole x-layers of the cube so it stays a block
    side = round(len(x) ** (1.0 / dim))
    if side ** dim == len(x) and n_sub < len(x):
        layers = max(1, n_sub // side ** (dim - 1))
        xs = np.ascontiguousarray(x[:layers * side ** (dim - 1)])
    else:
        xs = np.ascontiguousarray(x[:n_sub])
    sl = SlabSimulation(xs, model, res, device=local, rebalance_every=2)
    sl.advance(steps)
    got = sl.particles(dst=0)           # raises if a particle was lost or duplicated
    n_sum = int(_all_sum(sl.num_local(), local))
    out = {"particles": int(len(xs)), "steps": steps, "sum_num_local": n_sum, "migrated_records": int(_all_sum(sl.migrated, local))}
    del sl
    if rank == 0:
        twin = _s.MPMSimulation(xs, model, res, device=local)
        twin.advance(steps)
        ref = twin.particles()
        del twin
        vmax = max(1.0, float(np.abs(ref["v"]).max()))
        dv = np.abs(got["v"] - ref["v"]).max(axis=1)
        out.update(max_dx=float(np.abs(got["x"] - ref["x"]).max()), median_dv_over_vmax=float(np.median(dv)) / vmax,
                   p999_dv_over_vmax=float(np.quantile(dv, 0.999)) / vmax, max_dJp=float(np.abs(got["Jp"] - ref["Jp"]).max()),
                   max_dF=float(np.abs(got["F"] - ref["F"]).max()))
        ok = (n_sum == len(xs) and out["max_dx"] <= 3e-5 and out["median_dv_over_vmax"] <= 1e-3 and
              out["p999_dv_over_vmax"] <= 5e-3)
        out["ok"] = bool(ok)
    dist.barrier()
    return out


def _all_sum(v, local):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(v)], device=torch.device("cuda", local), dtype=torch.float64)
    dist.all_reduce(t)
    return float(t[0])


def bench_slabs(args, x, model, res, desc, rank, world, local):
    """Strong-scaling bench of one scene across `world` GPUs.  Device time = CUDA events around the step loop
    on each rank, bracketed by barrier + synchronize, MAX over ranks."""
    import torch
    import torch.distributed as dist
    from bench import ClockSampler, bench_config, measured_peak_gbs  # noqa: F401

    torch.cuda.set_device(local)
    verification = None if getattr(args, "no_verify", False) else verify_slabs(x, model, res, rank, world, local)
    sim = SlabSimulation(x, model, res, device=local, rebalance_every=args.rebalance_every)
    n_total = len(x)

    def sync():
        sim.synchronize()
        torch.cuda.synchronize()
        dist.barrier()

    sim.advance(args.warmup)
    sync()
    l0 = sim.engine.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        sync()
        e0.record()
        sim.advance(args.steps)
        e1.record()
        sync()
        ms = e0.elapsed_time(e1)
    launches = sim.engine.launch_count() - l0
    t = torch.tensor([ms, float(launches), float(sim.num_local()), float(sim.migrated)], device="cuda", dtype=torch.float64)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone()
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    tmin = t.clone()
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    ms = float(tmax[0])
    clocks = clk.summary()
    # per-phase device times of every slab (separate pass: the events serialise the host)
    sim.engine.sim.timing_enable(True)
    sim.engine.sim.timing_read(reset=True)
    sim.advance(min(args.steps, 10))
    sync()
    tm = sim.engine.sim.timing_read(reset=True)
    sim.engine.sim.timing_enable(False)
    ph = torch.tensor([tm[k] / max(1, tm["steps"]) for k in ("sort", "p2g", "grid", "g2p")] + [float(sim.num_local())],
                      device="cuda", dtype=torch.float64)
    ph_all = [torch.zeros_like(ph) for _ in range(world)]
    dist.all_gather(ph_all, ph)
    # end-to-end: the same steps including a gather of the positions to the host of every rank
    e2e_steps = max(3, min(args.steps, 10))
    host = sim.engine.pinned_slot_buffers(sim.capacity)
    sim.advance(1)
    local_state = sim.engine.download_slots(host, compact=False)
    sync()
    import time
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sim.advance(1)
        local_state = sim.engine.download_slots(host, compact=False)
    sync()
    t_e2e = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    d2h = sum(v.nbytes for v in local_state.values())
    # ---- bookkeeping checks on the state the timed steps left behind: every particle exactly once -----------------
    ids = local_state["ids"]
    live = ids[ids != 0xFFFFFFFF].astype(np.uint64)
    chk = torch.tensor([float(len(live)), float(live.sum()), float((live * live % np.uint64(1000003)).sum())],
                       device="cuda", dtype=torch.float64)
    dist.all_reduce(chk)
    idx = np.arange(n_total, dtype=np.uint64)
    want = [float(n_total), float(idx.sum()), float((idx * idx % np.uint64(1000003)).sum())]
    nl = torch.tensor([float(sim.num_local())], device="cuda", dtype=torch.float64)
    dist.all_reduce(nl)
    checks = {"sum_num_local": int(nl[0]), "live_slots": int(chk[0]), "id_sum_ok": bool(chk[1] == want[1]),
              "id_square_hash_ok": bool(chk[2] == want[2]),
              "ok": bool(int(nl[0]) == n_total and int(chk[0]) == n_total and chk[1] == want[1] and chk[2] == want[2])}
    if rank != 0:
        return None
    if not checks["ok"] or (verification is not None and not verification.get("ok", False)):
        raise RuntimeError(f"slab bench: correctness checks failed: {checks} {verification}")
    dim = x.shape[1]
    return {
        "metric": "particle-steps/s", "value": n_total * args.steps / (ms * 1e-3), "unit": "particle-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, desc, n_total, res, dim, model),
        "run": {"partition": f"{world} x-slabs by particle count", "bounds": sim.bounds,
                "rebalance_every": args.rebalance_every, "particles_per_rank_min_max": [int(tmin[2]), int(tmax[2])],
                "migrated_records_total": int(tsum[3])},
        "checks": checks, "verification": verification,
        "clocks": clocks, "gpu_launches": int(tsum[1]),
        "e2e": {"value": n_total * e2e_steps / float(t_e2e[0]), "unit": "particle-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                "what": "per step: slab advance(1) + download of the rank's particle state (slot order + global ids) "
                        "into pinned host memory (no per-step upload: a slab's particle set changes by migration)"},
        "roofline": _slab_roofline(ph_all, dim), "cpu_baseline": None,
    }


def _slab_roofline(ph_all, dim):
    """Roofline of the dominant kernel on the slowest slab: algorithmic bytes of that slab's particles over the
    kernel's mean device time there, against the measured HBM peak of ONE GPU."""
    from bench import ALGO_BYTES, measured_peak_gbs, ncu_traffic
    peak, src = measured_peak_gbs()
    rows = [[float(v) for v in t.tolist()] for t in ph_all]
    r = max(range(len(rows)), key=lambda i: rows[i][1] + rows[i][3])
    sort_ms, p2g_ms, grid_ms, g2p_ms, n_local = rows[r]
    dom = "g2p" if g2p_ms >= p2g_ms else "p2g"
    dom_ms = g2p_ms if dom == "g2p" else p2g_ms
    ach = ALGO_BYTES[dim][dom] * n_local / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    return {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": ncu_traffic(dom, n_local), "peak_source": src, "rank": r, "particles_on_rank": int(n_local),
            "algorithmic_bytes_per_particle": ALGO_BYTES[dim][dom],
            "phase_ms_per_rank": [dict(sort_ms=a, p2g_ms=b, grid_ms=c, g2p_ms=d, particles=int(e)) for a, b, c, d, e in rows]}
