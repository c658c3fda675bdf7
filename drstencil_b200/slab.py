"""Slab decomposition of a 3D grid along k (the slowest axis) over the GPUs of one box, one process
per GPU (torchrun), with nearest-neighbour halo exchange after every sweep.

The reference is single-GPU; this is the engine's extension for BASELINE.json
config c5.  Semantics are those of ONE global grid swept by the reference's schedule: the global
Halo-wide ring stays frozen, every rank owns planes [lo, hi) and keeps `ghost` = Halo copies of
its neighbours' boundary planes on each side.

Two exchange paths:
  "p2p"   the sweep kernel itself stores its boundary planes a second time, straight into the
          neighbour's ghost planes over NVLink (buffers mapped with CUDA IPC, drs_plan_set_peers);
          ranks then only trade a step flag (drs_signal_peers / drs_wait_flags) -- no collective,
          no extra copy kernel, and the transfer overlaps the rest of the sweep;
  "nccl"  plain torch.distributed isend/irecv of the boundary planes after each sweep (NCCL on GPUs,
          gloo on CPU -- the path the CPU tests drive with a stand-in sweep).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Callable, List, Optional


def split(n: int, parts: int):
    """[(lo, hi)] -- contiguous, as even as possible."""
    return [(r * n // parts, (r + 1) * n // parts) for r in range(parts)]


@dataclass
class SlabGeometry:
    """Where one rank's slab sits in the global grid (all plane indices)."""
    g_slow: int      # global planes
    world: int
    rank: int
    ghost: int       # Halo = step * order: ghost planes per side

    def __post_init__(self):
        b = split(self.g_slow, self.world)
        self.bounds = b
        self.lo, self.hi = b[self.rank]
        if min(h - l for l, h in b) < self.ghost:
            raise ValueError("slab thinner than the halo: %d planes over %d ranks with Halo %d"
                             % (self.g_slow, self.world, self.ghost))
        self.origin = self.lo - self.ghost                 # global index of local plane 0
        self.local_planes = self.hi - self.lo + 2 * self.ghost
        self.lower = self.rank - 1 if self.rank > 0 else None
        self.upper = self.rank + 1 if self.rank + 1 < self.world else None
        # output range in local indices (the global ring stays frozen)
        self.out_lo = max(self.lo, self.ghost) - self.origin
        self.out_hi = min(self.hi, self.g_slow - self.ghost) - self.origin

    def local(self, g: int) -> int:
        return g - self.origin

    # planes (local indices) this rank sends to / receives from each neighbour after a sweep
    def send_lower(self):
        return (self.local(self.lo), self.local(self.lo + self.ghost))

    def send_upper(self):
        return (self.local(self.hi - self.ghost), self.local(self.hi))

    def recv_lower(self):
        return (self.local(self.lo - self.ghost), self.local(self.lo))

    def recv_upper(self):
        return (self.local(self.hi), self.local(self.hi + self.ghost))


def halo_exchange(buf, geom: SlabGeometry, group=None) -> None:
    """Refreshes the ghost planes of `buf` (a [local_planes, M, N] tensor) from the neighbours'
    boundary planes with point-to-point sends/receives."""
    import torch.distributed as dist
    ops = []
    if geom.lower is not None:
        a, b = geom.send_lower()
        ops.append(dist.P2POp(dist.isend, buf[a:b], geom.lower, group))
        a, b = geom.recv_lower()
        ops.append(dist.P2POp(dist.irecv, buf[a:b], geom.lower, group))
    if geom.upper is not None:
        a, b = geom.send_upper()
        ops.append(dist.P2POp(dist.isend, buf[a:b], geom.upper, group))
        a, b = geom.recv_upper()
        ops.append(dist.P2POp(dist.irecv, buf[a:b], geom.upper, group))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()


class SlabRunner:
    """The reference's ping-pong schedule over a slab.  `sweep(src, dst)` advances the local
    output range of `dst` from `src`; `after(dst, s)` makes the neighbours' ghosts current."""

    def __init__(self, geom: SlabGeometry, bufs, sweep: Callable, after: Callable, before: Optional[Callable] = None):
        self.geom, self.bufs, self.sweep, self.after, self.before = geom, bufs, sweep, after, before
        self.sweeps_done = 0

    def run(self, timesteps: int, step: int) -> int:
        n = 0
        t = 0
        while t < timesteps:
            for _ in range(2):
                s = self.sweeps_done
                src, dst = self.bufs[s & 1], self.bufs[(s & 1) ^ 1]
                if self.before:
                    self.before(s)
                self.sweep(src, dst)
                self.after(dst, s)
                self.sweeps_done += 1
                n += 1
            t += 2 * step
        return n


# ---------------------------------------------------------------------------------------------
# GPU wiring
# ---------------------------------------------------------------------------------------------

class _DevArray:
    """A cudaMalloc allocation (whole allocation, so it can be exported with CUDA IPC) viewed as a
    torch tensor through __cuda_array_interface__."""

    def __init__(self, shape, np_typestr: str, itemsize: int):
        from . import lib, _check
        n = 1
        for d in shape:
            n *= d
        p = ctypes.c_void_p()
        _check(lib().drs_device_malloc(n * itemsize, ctypes.byref(p)))
        self.ptr = p.value
        self.nbytes = n * itemsize
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": np_typestr, "data": (self.ptr, False),
                                         "version": 2, "strides": None}

    def tensor(self):
        import torch
        return torch.as_tensor(self, device="cuda")

    def free(self):
        from . import lib
        if self.ptr:
            lib().drs_device_free(self.ptr)
            self.ptr = 0


def _ipc_export(ptr: int) -> bytes:
    from . import lib, _check
    h = ctypes.create_string_buffer(64)
    _check(lib().drs_ipc_export(ptr, h))
    return h.raw


def _ipc_import(handle: bytes) -> int:
    from . import lib, _check
    p = ctypes.c_void_p()
    _check(lib().drs_ipc_import(handle, ctypes.byref(p)))
    return p.value


class GpuSlab:
    """One rank's share of a slab-decomposed run on its GPU."""

    def __init__(self, stc_path: str, knobs, rank: int, world: int, halo: str = "p2p", group=None,
                 global_shape=None):
        import torch
        import torch.distributed as dist
        from . import Plan, Stencil, F32
        self.rank, self.world, self.mode, self.group = rank, world, halo, group
        st = Stencil.from_file(stc_path)
        if st.dim != 3:
            raise ValueError("slab decomposition is implemented for 3D grids")
        if global_shape is not None:
            st.set_size(global_shape)
        self.global_shape = tuple(st.shape)
        L, M, N = self.global_shape
        probe = Plan(st, knobs)                      # halo of the (possibly multi-step) sweep
        ghost = probe.halo
        self.step = knobs.step
        self.geom = SlabGeometry(L, world, rank, ghost)
        st.set_size((self.geom.local_planes, M, N))
        self.plan = Plan(st, knobs)
        self.plan.set_slab(L, self.geom.lo, self.geom.hi)
        self.dtype = torch.float32 if knobs.dtype == F32 else torch.float64
        typestr, isz = ("<f4", 4) if knobs.dtype == F32 else ("<f8", 8)
        shape = (self.geom.local_planes, M, N)
        self._raw = [_DevArray(shape, typestr, isz), _DevArray(shape, typestr, isz)]
        self.bufs = [r.tensor() for r in self._raw]
        for b in self.bufs:
            b.zero_()
        self._peer_ptrs: List[int] = []
        if halo == "p2p" and world > 1:
            self._flags = _DevArray((2,), "<i8", 8)
            self._flags.tensor().zero_()
            mine = [_ipc_export(r.ptr) for r in self._raw] + [_ipc_export(self._flags.ptr)]
            everyone = [None] * world
            dist.all_gather_object(everyone, mine, group=group)
            lower = upper = None
            g = self.geom
            if g.lower is not None:
                lower = [_ipc_import(h) for h in everyone[g.lower]]
                self._peer_ptrs += lower
            if g.upper is not None:
                upper = [_ipc_import(h) for h in everyone[g.upper]]
                self._peer_ptrs += upper
            self.plan.set_peers([r.ptr for r in self._raw], lower[:2] if lower else [0, 0],
                                upper[:2] if upper else [0, 0],
                                g.bounds[g.lower][0] if lower else 0, g.bounds[g.upper][0] if upper else 0)
            # a rank writes slot 1 of its lower neighbour's flags and slot 0 of its upper neighbour's
            self._lower_flag = lower[2] + 8 if lower else 0
            self._upper_flag = upper[2] if upper else 0
            dist.barrier(group=group)
        self.runner = SlabRunner(self.geom, self.bufs, self._sweep, self._after, self._before)

    # -- schedule hooks --
    def _before(self, s: int) -> None:
        if self.mode == "p2p" and self.world > 1 and s > 0:
            g = self.geom
            self.plan.wait_flags(self._flags.ptr, g.lower is not None, g.upper is not None, s)

    def _sweep(self, src, dst) -> None:
        self.plan.sweep(src, dst)

    def _after(self, dst, s: int) -> None:
        if self.world == 1:
            return
        if self.mode == "p2p":
            self.plan.signal_peers(self._lower_flag, self._upper_flag, s + 1)
        else:
            halo_exchange(dst, self.geom, self.group)

    # -- data --
    def fill(self, plane_fn: Callable) -> None:
        """A[local plane] = plane_fn(global plane index) for every in-grid plane this rank holds
        (owned and ghost); B = 0.  Resets the schedule."""
        import torch.distributed as dist
        g = self.geom
        self.bufs[0].zero_()
        self.bufs[1].zero_()
        for zl in range(g.local_planes):
            zg = g.origin + zl
            if 0 <= zg < g.g_slow:
                self.bufs[0][zl].copy_(plane_fn(zg))
        if self.mode == "p2p" and self.world > 1:
            self.plan.sync_check()
            self._flags.tensor().zero_()
            import torch
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
        self.runner.sweeps_done = 0

    def run(self, timesteps: int) -> int:
        return self.runner.run(timesteps, self.step)

    def run_host(self, h_own, timesteps: int) -> float:
        """EXPERIMENTAL (not yet validated on GPUs): the reference schedule on this rank's HOST slab `h_own`
        (own planes only, pinned), uploads / sweeps / downloads overlapped by time-skewed blocks, faces in
        lockstep with the neighbours (drs_run_host_slab).  Needs the "p2p" halo mode and a GpuSlab of its
        own: run() and run_host() number the step flags differently and must not be mixed on one object.
        Returns device ms."""
        import torch
        import torch.distributed as dist
        if self.mode != "p2p" or self.world < 2:
            raise ValueError("run_host needs the p2p halo mode and at least two ranks (one GPU: Plan.run_host)")
        if self.runner.sweeps_done:
            raise ValueError("this GpuSlab has been used with run(); create a separate one for run_host()")
        n = 0
        t = 0
        while t < timesteps:
            n += 2
            t += 2 * self.step
        self.plan.sync_check()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)               # nobody still reads ghosts of the previous call
        base = getattr(self, "_flag_base", 0)
        ms = self.plan.run_host_slab(h_own, timesteps, self.rank % 2 == 1, self._flags.ptr, self._lower_flag,
                                     self._upper_flag, base)
        self._flag_base = base + n + 2
        return ms

    def owned(self, which: int = 0):
        """The rank's owned planes of buffer `which` (0 = A, where the result lands)."""
        g = self.geom
        return self.bufs[which][g.ghost:g.ghost + (g.hi - g.lo)]

    def close(self) -> None:
        from . import lib
        import torch
        torch.cuda.synchronize()
        for p in self._peer_ptrs:
            lib().drs_ipc_close(p)
        self._peer_ptrs = []


def bench_slab(args, rank, world, workload, peak_info, timed_start=None, timed_end=None):
    """bench.py's N > 1 leg: c5 (3d7pt_star fp64 1536^3) over `world` GPUs, strong scaling.
    `timed_start` / `timed_end` bracket the timed region (clock sampling)."""
    import time
    import torch
    import torch.distributed as dist
    from . import sweep_count
    from .presets import PRESETS
    preset, timesteps, desc = workload
    path, kn = PRESETS[preset]
    if getattr(args, "depth", 1) > 1:
        from . import Knobs
        kn = Knobs(step=args.depth)
        desc += " [temporal depth %d]" % args.depth
    slab = GpuSlab(path, kn, rank, world, halo=args.halo)
    L, M, N = slab.global_shape
    g = torch.Generator(device="cuda")
    dtype = slab.dtype

    def plane(zg):
        g.manual_seed(1234 + zg)
        return torch.rand((M, N), dtype=dtype, device="cuda", generator=g) * 1e-100

    slab.fill(plane)
    info = slab.plan.info
    for _ in range(max(3, args.warmup)):
        slab.run(timesteps)
    slab.plan.sync_check()
    dist.barrier()
    torch.cuda.synchronize()
    l0 = slab.plan.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if timed_start:
        timed_start()
    e0.record()
    for _ in range(args.steps):
        slab.run(timesteps)
    e1.record()
    slab.plan.sync_check()
    clocks = timed_end() if timed_end else None
    secs = e0.elapsed_time(e1) * 1e-3
    t = torch.tensor([secs], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    secs = float(t)
    launches = slab.plan.launch_count - l0
    H = info.halo
    sweeps = sweep_count(timesteps, kn.step)
    upd = (L - 2 * H) * (M - 2 * H) * (N - 2 * H) * sweeps * kn.step
    value = upd * args.steps / secs / 1e9
    peak, peak_src = peak_info
    esize = 8 if slab.dtype == torch.float64 else 4
    geom = slab.geom
    local_bytes = (geom.hi - geom.lo) * M * N * 2 * esize        # algorithmic bytes of this rank's launch
    sweep_launches = sweeps * args.steps
    ach = local_bytes / (secs / sweep_launches) / 1e9
    halo_bytes = 2 * geom.ghost * M * N * esize                  # pushed per sweep by an interior rank
    line = {
        "metric": "GStencil/s", "value": value, "unit": "GStencil/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64" if esize == 8 else "f32", "data": "synthetic",
        "config": {"workload": "%s: %s" % (preset, desc), "grid": [L, M, N], "timesteps_per_step": timesteps,
                   "decomposition": "k-slabs, %d planes per GPU + %d ghost planes per side" % (geom.hi - geom.lo, geom.ghost),
                   "halo_exchange": "fused NVLink peer stores from the sweep kernel + step flags" if args.halo == "p2p"
                   else "NCCL isend/irecv after each sweep",
                   "halo_bytes_per_sweep_per_gpu": halo_bytes, "kernel": info.kernel_name,
                   "l2": "inputs larger than L2 (%.1f GiB per rank per sweep)" % (local_bytes / 2 ** 30)},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                     "peak_source": peak_src, "kernel": info.kernel_name, "per": "GPU (rank 0's slab)",
                     "algorithmic_bytes_per_launch": local_bytes, "launch_ms": secs / sweep_launches * 1e3},
        "clocks": clocks,
    }
    # e2e: pinned host slab -> device, the schedule, result back (every step)
    own = slab.owned(0)
    h = torch.empty(own.shape, dtype=own.dtype, pin_memory=True)
    h.fill_(0.5e-100)

    def e2e_step():
        own.copy_(h, non_blocking=True)
        if world > 1:
            halo_exchange(slab.bufs[0], geom, None) if args.halo == "nccl" else _p2p_refresh(slab)
        slab.run(timesteps)
        h.copy_(own, non_blocking=True)

    e2e_steps = 2
    e2e_step()                  # untimed warm-up: the first ghost refresh sets up the NCCL point-to-point channels
    h.fill_(0.5e-100)
    slab.plan.sync_check()
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    slab.plan.sync_check()
    es = e0.elapsed_time(e1) * 1e-3
    t = torch.tensor([es], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    es = float(t)
    line["e2e"] = {"value": upd * e2e_steps / es / 1e9, "unit": "GStencil/s", "h2d_bytes_per_step": own.numel() * esize * world,
                   "d2h_bytes_per_step": own.numel() * esize * world, "steps": e2e_steps, "ms_per_step": es / e2e_steps * 1e3,
                   "api": "GpuSlab.run on pinned host slabs (one process per GPU)"}
    slab.close()
    for r in slab._raw:
        r.free()
    del slab
    torch.cuda.empty_cache()
    # extra evidence in the same run: the same grid with in-kernel temporal depth 2 (fused halo push of
    # two ghost planes per side); the contract value above stays the bit-exact depth-1 sweep
    if getattr(args, "depth", 1) == 1 and not getattr(args, "no_extras", False):
        try:
            from . import Knobs
            kn2 = Knobs(step=2)
            s2 = GpuSlab(path, kn2, rank, world, halo=args.halo)
            s2.fill(plane)
            s2.run(timesteps)
            s2.plan.sync_check()
            dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                s2.run(timesteps)
            e1.record()
            s2.plan.sync_check()
            t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            H2 = s2.plan.info.halo
            upd2 = (L - 2 * H2) * (M - 2 * H2) * (N - 2 * H2) * sweep_count(timesteps, 2) * 2
            line["temporal_fused"] = {"depth": 2, "value": upd2 * 3 / float(t) / 1e9, "unit": "GStencil/s",
                                      "ms_per_step": float(t) / 3 * 1e3, "kernel": s2.plan.info.kernel_name,
                                      "parity": "<= 1e-12 relative vs the composed operator (tests), bit-identical to the single-GPU run"}
            s2.close()
        except Exception as e:   # extra only
            line["temporal_fused"] = {"error": str(e)[:200]}
    return line


def _p2p_refresh(slab: "GpuSlab") -> None:
    """After new host data was uploaded into the owned planes, rebuild the neighbours' ghosts with
    a plain exchange (the fused push only covers planes a sweep has just produced)."""
    halo_exchange(slab.bufs[0], slab.geom, slab.group)
