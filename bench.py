#!/usr/bin/env python
"""bench.py -- the driver's contract benchmark.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cX]

Metric (BASELINE.json): GStencil/s and achieved HBM GB/s (% of roofline) per stencil.
A "step" is one run of the reference's emitted host loop (codegen_2d.hpp:610-613) over one
synthetic grid: `for (t = 0; t < iterations; t += 2*step) { sweep(A,B); sweep(B,A); }`.

  Workload, every N: c5 = 3d7pt_star fp64 1536^3, 100 timesteps per step -- the one BASELINE.json
          configuration defined at 1/2/4/8 GPUs (it fits one B200: 2 x 27 GiB), so that the per-N
          values form one strong-scaling series.  N = 1 sweeps the whole grid on one GPU; N > 1
          slab-decomposes it along k with the fused NVLink halo push.
  N = 1   additionally reports every other configuration (c1-c4) in `per_config`, each with its
          own roofline fraction -- c2 (2d9pt_box fp64 16384^2, temporal depth 4) is the
          temporally fused one -- and the CPU baseline.

`value`     device-resident throughput (inputs already in HBM), CUDA events, max over ranks.
`e2e`       the same step through the host-buffer C-ABI call drs_run_host (pinned host memory:
            H2D of the grid, the schedule, D2H of the result inside the timed region).
`roofline`  dominant kernel: algorithmic bytes (2*sizeof(T) per grid point per launch)
            / mean launch time, against MEASURED_PEAKS.json hbm_gbs (burst copy figure).
`cpu_baseline`  the CPU oracle (a port: the reference has no CPU loop) on all host cores, on a
            bounded sample of the same workload.
`--impl reference`  the same oracle port, timed as the reference arm (rank 0 only).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (preset, timesteps per step, description)
    "c1": ("c1", 10, "2d5pt_star fp64 4096^2, 10 timesteps"),
    "c2": ("c2", 128, "2d9pt_box fp64 16384^2, temporal fusion depth 4, 128 timesteps"),
    "c3": ("c3", 8, "2d25pt_box fp32 16384^2, 8 timesteps"),
    "c4": ("c4", 16, "3d7pt_star fp64 768^3, k-streaming, 16 timesteps"),
    "c5": ("c5", 100, "3d7pt_star fp64 1536^3 slab-decomposed along k, 100 timesteps"),
    "c4t2": ("c4t2", 16, "EXTRA (not in BASELINE.json): 3d7pt_star fp64 768^3 with in-kernel temporal depth 2, 16 timesteps"),
    "c5t2": ("c5t2", 20, "EXTRA: c5 (3d7pt_star fp64 1536^3) with in-kernel temporal depth 2, 20 timesteps"),
    "c1t2": ("c1t2", 20, "EXTRA: c1 (2d5pt_star fp64 4096^2) with in-kernel temporal depth 2, 20 timesteps"),
    "c3t2": ("c3t2", 8, "EXTRA: c3 (2d25pt_box fp32 16384^2) with in-kernel temporal depth 2, 8 timesteps"),
}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def interior_points(shape, halo):
    n = 1
    for d in shape:
        n *= max(0, d - 2 * halo)
    return n


def all_points(shape):
    n = 1
    for d in shape:
        n *= d
    return n


def time_steps(plan, A, B, timesteps, steps, warmup):
    """K timed steps of the emitted host loop on device-resident buffers -> (seconds, launches)."""
    import torch
    for _ in range(warmup):
        plan.run(A, B, timesteps)
    plan.sync_check()
    l0 = plan.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        plan.run(A, B, timesteps)
    e1.record()
    plan.sync_check()
    return e0.elapsed_time(e1) * 1e-3, plan.launch_count - l0


class CpuSample:
    """The oracle port on all host cores over a bounded sample of a workload: the composed operator
    of the workload (gold restatement) swept over a sub-grid with the workload's full plane / row
    size, so that the cache behaviour per plane is the real one; only the slow axis is cut."""

    def __init__(self, workload):
        import numpy as np
        from oracle import oracle
        import drstencil_b200 as drs
        from drstencil_b200.presets import PRESETS
        self.oracle = oracle
        self.workload = workload
        path, kn = PRESETS[WORKLOADS[workload][0]]
        is3d = os.path.basename(path)[3:].startswith("3d")
        s = oracle.parse_stc(path, is3d)
        self.step = kn.step
        pts = oracle.compose(s.points, self.step)
        self.offs, self.coefs = oracle.terms(pts)
        self.halo, _ = oracle.order_dist(pts, s.dim)
        dtype = np.float32 if kn.dtype == drs.F32 else np.float64
        # ~3.5 GiB per array at most: 192 planes of c5, 4096 rows of the 16384-wide 2D grids
        self.shape = (min(s.L, 192), s.M, s.N) if is3d else (min(s.M, 4096), s.N)
        self.a = oracle.lcg_array(self.shape, dtype, 1)
        self.b = oracle.lcg_array(self.shape, dtype, 2)
        oracle.sweep(self.a, self.b, self.offs, self.coefs, self.halo)     # warm-up / first touch
        self.cores = oracle.lib().drs_oracle_threads()

    def run(self, budget_s):
        """Sweeps for about budget_s seconds -> (GStencil/s over the sample, seconds, sweeps)."""
        used, sweeps = 0.0, 0
        while used < budget_s or sweeps < 2:
            t0 = time.perf_counter()
            self.oracle.sweep(self.a, self.b, self.offs, self.coefs, self.halo)
            used += time.perf_counter() - t0
            sweeps += 1
            self.a, self.b = self.b, self.a
            if sweeps % 8 == 0:       # sum of coefficients > 1: keep the values finite
                self.a *= 1e-3
        return interior_points(self.shape, self.halo) * self.step * sweeps / used / 1e9, used, sweeps

    def describe(self, sweeps, seconds):
        return ("%s sub-grid %s (full-size planes, slow axis cut), composed %d-point operator (gold restatement, "
                "%d timesteps per sweep), %d sweeps in %.1f s" % (self.workload, "x".join(map(str, self.shape)),
                                                                   len(self.coefs), self.step, sweeps, seconds))


def cpu_baseline(workload, budget_s=12.0):
    """bench.py's cpu_baseline object: about budget_s seconds of CPU work."""
    cs = CpuSample(workload)
    gst, used, sweeps = cs.run(budget_s)
    return {"value": gst, "unit": "GStencil/s", "cores": cs.cores, "kind": "port",
            "sample": cs.describe(sweeps, used), "seconds_per_sweep": used / sweeps}


def reference_gpu_kernel(workload):
    """The reference's own emitted dr_ kernel (oracle/_ref, nvcc sm_100a) on this GPU, for context."""
    import ctypes
    case = {"c1": "full_c1", "c2": "full_c2", "c4": "full_c4"}.get(workload)
    meta_p = os.path.join(ROOT, "oracle", "_ref", "cases.json")
    if case is None or not os.path.exists(meta_p):
        return None
    meta = json.load(open(meta_p)).get(case)
    if not meta:
        return None
    try:
        lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", meta["so"]))
        lib.drs_ref_time.restype = ctypes.c_float
        lib.drs_ref_time.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
        ms = lib.drs_ref_time(1, 6, 2) / 6
        if ms <= 0:
            return None
        halo = meta["step"] * 1
        shape = (meta["L"], meta["M"], meta["N"]) if meta["is3d"] else (meta["M"], meta["N"])
        return {"value": interior_points(shape, halo) * meta["step"] / (ms * 1e-3) / 1e9, "unit": "GStencil/s",
                "ms_per_sweep": ms, "kernel": "dr_%s emitted by the reference generator (%s), nvcc sm_100a"
                                              % (meta["stencil"], " ".join(meta["options"]))}
    except Exception as e:   # context only
        return {"error": str(e)[:200]}


def run_single(args, rank, world):
    import torch
    import drstencil_b200 as drs
    from drstencil_b200.presets import PRESETS
    wl = args.workload or "c5"
    preset, timesteps, desc = WORKLOADS[wl]
    path, kn = PRESETS[preset]
    if wl == "c5" and args.depth > 1:
        kn = drs.Knobs(step=args.depth)
        desc += " [temporal depth %d]" % args.depth
    st = drs.Stencil.from_file(path)
    plan = drs.Plan(st, kn)
    shape = st.shape
    dtype = torch.float32 if kn.dtype == drs.F32 else torch.float64
    esize = 4 if kn.dtype == drs.F32 else 8
    g = torch.Generator(device="cuda").manual_seed(rank)
    A = torch.rand(shape, dtype=dtype, device="cuda", generator=g)
    if dtype == torch.float64:
        A.mul_(1e-200)   # sum of coefficients > 1: keeps (K + W) * timesteps updates inside fp64 range
    B = torch.zeros_like(A)
    sampler = ClockSampler(torch.cuda.current_device())
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    secs, launches = time_steps(plan, A, B, timesteps, args.steps, args.warmup)
    clocks = sampler.stop()
    info = plan.info
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([secs], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t)
    sweeps_per_step = drs.sweep_count(timesteps, kn.step)
    upd = interior_points(shape, info.halo) * sweeps_per_step * kn.step   # stencil updates per step
    value = world * upd * args.steps / secs / 1e9
    peak, peak_src = measured_peak()
    alg_bytes = all_points(shape) * 2 * esize
    ach = alg_bytes / (secs / launches) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        # keyed by workload (tools/make_profiles.py); --depth n swaps the kernel, for which there is no capture
        traffic = json.load(open(tp)).get(wl if kn.step == PRESETS[preset][1].step else "")
    line = {
        "metric": "GStencil/s", "value": value, "unit": "GStencil/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if wl == "c5" else "weak", "vs_baseline": None,
        "dtype": "f32" if kn.dtype == drs.F32 else "f64", "data": "synthetic",
        "config": {"workload": "%s: %s" % (wl, desc), "grid": list(shape), "timesteps_per_step": timesteps,
                   "sweeps_per_step": sweeps_per_step, "temporal_depth": kn.step, "kernel": info.kernel_name,
                   "tile": {"warps_per_cta": info.warps_per_cta, "tile_x": info.tile_x, "chunk": info.chunk,
                            "stages": info.stages, "regs": info.regs_per_thread, "smem": info.smem_bytes},
                   "l2": "inputs larger than L2 (2 x %.2f GiB per sweep vs 126 MB)" % (alg_bytes / 2 / 2 ** 30),
                   "replicas": world if world > 1 else None},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": info.kernel_name,
                     "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": secs / launches * 1e3,
                     "gstencil_roofline": peak / (2 * esize) * kn.step},
        "clocks": clocks,
    }
    # ---- e2e: the host-buffer entry point (H2D + schedule + D2H in the timed region) ----
    if rank == 0 or world > 1:
        del A, B
        torch.cuda.empty_cache()
        pinned = True
        try:
            hA = torch.empty(shape, dtype=dtype, pin_memory=True)
        except RuntimeError:
            pinned = False
            hA = torch.empty(shape, dtype=dtype)
        hA.fill_(0.5e-100)
        e2e_steps = 2 if all_points(shape) * esize > 8 * 2 ** 30 else max(2, min(args.steps, 5))
        plan.run_host(hA, None, timesteps)               # warm (allocates the device pair)
        dev_ms = 0.0
        for _ in range(e2e_steps):
            dev_ms += plan.run_host(hA, None, timesteps)
        e_secs = dev_ms * 1e-3
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([e_secs], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_secs = float(t)
        nbytes = all_points(shape) * esize
        # the same call with the overlap switched off (copy, sweep, copy in sequence), for reference
        plan.set_host_block(-1)
        plain_ms = plan.run_host(hA, None, timesteps)
        plan.set_host_block(0)
        line["e2e"] = {"value": world * upd * e2e_steps / e_secs / 1e9, "unit": "GStencil/s",
                       "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "steps": e2e_steps,
                       "ms_per_step": e_secs / e2e_steps * 1e3,
                       "api": "drs_run_host (C ABI): %s host grid -> H2D, schedule, D2H of the result; the three phases "
                              "overlapped by time-skewed blocks along the slow axis (bit-identical to the plain sequence)"
                              % ("pinned" if pinned else "pageable"),
                       "plain_sequence_ms_per_step": plain_ms}
        del hA
    return line, plan


def per_config(args):
    """The other single-GPU BASELINE configs, device-resident, a few sweeps each."""
    import torch
    import drstencil_b200 as drs
    from drstencil_b200.presets import PRESETS
    peak, _ = measured_peak()
    out = []
    for wl in ("c1", "c2", "c3", "c4", "c1t2", "c3t2", "c4t2", "c5t2"):
        preset, timesteps, desc = WORKLOADS[wl]
        path, kn = PRESETS[preset]
        st = drs.Stencil.from_file(path)
        shape = st.shape
        esize = 4 if kn.dtype == drs.F32 else 8
        if 2 * all_points(shape) * esize > 0.8 * torch.cuda.get_device_properties(0).total_memory:
            continue
        try:
            plan = drs.Plan(st, kn)
            dtype = torch.float32 if kn.dtype == drs.F32 else torch.float64
            A = torch.rand(shape, dtype=dtype, device="cuda")
            B = torch.zeros_like(A)
            ts = timesteps if wl != "c2" else 32
            k = 5 if wl == "c1" else 2
            if dtype == torch.float64:
                A.mul_(1e-100)
            secs, launches = time_steps(plan, A, B, ts, k, 2)
            info = plan.info
            upd = interior_points(shape, info.halo) * drs.sweep_count(ts, kn.step) * kn.step * k
            ach = all_points(shape) * 2 * esize / (secs / launches) / 1e9
            # what a plain device copy of the same array reaches (small grids cannot reach the big-copy peak)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(3):
                B.copy_(A)
            e0.record()
            for _ in range(10):
                B.copy_(A)
            e1.record()
            torch.cuda.synchronize()
            copy_gbs = all_points(shape) * 2 * esize / (e0.elapsed_time(e1) * 1e-4) / 1e9
            out.append({"workload": "%s: %s" % (wl, desc), "value": upd / secs / 1e9, "unit": "GStencil/s",
                        "temporal_depth": kn.step, "kernel": info.kernel_name, "launch_ms": secs / launches * 1e3,
                        "gpu_launches": launches, "n_gpus": 1,
                        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                     "gstencil_roofline": peak / (2 * esize) * kn.step,
                                     "frac_of_single_step_gstencil_roofline": (upd / secs / 1e9) / (peak / (2 * esize)),
                                     "torch_copy_same_array_gbs": copy_gbs,
                                     "traffic": (json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(wl)
                                                 if os.path.exists(os.path.join(ROOT, "profiles", "traffic.json")) else None)}})
            del A, B, plan
            torch.cuda.empty_cache()
        except Exception as e:
            out.append({"workload": wl, "error": str(e)[:200]})
    return out


def run_reference(args, rank):
    """Reference arm: the CPU oracle port (the reference ships no CPU loop and its emitted CUDA
    is not a CPU implementation), all host threads; each step is a bounded sample (~3 s of sweeps)
    of the workload, W warm-up steps, K timed steps."""
    if rank != 0:
        return None
    wl = args.workload or "c5"
    cs = CpuSample(wl)
    for _ in range(args.warmup):
        cs.run(3.0)
    vals, secs, sweeps = [], 0.0, 0
    for _ in range(max(1, args.steps)):
        v, t, n = cs.run(3.0)
        vals.append(v)
        secs += t
        sweeps += n
    v = interior_points(cs.shape, cs.halo) * cs.step * sweeps / secs / 1e9
    base = {"value": v, "unit": "GStencil/s", "cores": cs.cores, "kind": "port", "sample": cs.describe(sweeps, secs),
            "seconds_per_sweep": secs / sweeps, "per_step": vals}
    preset, timesteps, desc = WORKLOADS[wl]
    return {"impl": "reference", "metric": "GStencil/s", "value": v, "unit": "GStencil/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / max(1, args.steps) * 1e3,
            "higher_is_better": True, "scaling": "strong" if wl == "c5" else "weak", "vs_baseline": None,
            "dtype": "f64" if wl not in ("c3", "c3t2") else "f32", "data": "synthetic",
            "config": {"workload": "%s: %s" % (wl, desc), "note": "CPU port of the reference's gold expression "
                       "(the reference has no CPU implementation); each step = one bounded sample of ~3 s"},
            "cpu_baseline": base, "gpu_launches": 0,
            "e2e": {"value": v, "unit": "GStencil/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--no-extras", action="store_true", help="skip per_config / cpu_baseline (profiling runs)")
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"], help="N > 1: halo exchange path")
    ap.add_argument("--depth", type=int, default=1, help="c5 only: in-kernel temporal depth (extra evidence; "
                    "the contract line is depth 1, bit-exact)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line (the JSON): libraries that write to fd 1 from C (NCCL prints its
    # version there) are sent to stderr instead
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    if args.impl == "reference":
        line = run_reference(args, rank)
        if line is not None:
            print(json.dumps(line), file=json_out, flush=True)
        return

    import torch
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device -- the engine has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world > 1 and (args.workload in (None, "c5")):   # world == 1 sweeps the whole grid on one GPU
        from drstencil_b200 import slab
        sampler = ClockSampler(torch.cuda.current_device())
        line = slab.bench_slab(args, rank, world, WORKLOADS["c5"], measured_peak(), sampler.start, sampler.stop)
    else:
        line, plan = run_single(args, rank, world)
        del plan                     # frees the device pair drs_run_host allocated
        torch.cuda.empty_cache()
    if rank == 0:
        if world == 1 and not args.no_extras:
            line["per_config"] = per_config(args)
            line["cpu_baseline"] = cpu_baseline(args.workload or "c5")
            refs = {w: reference_gpu_kernel(w) for w in ("c1", "c2", "c4")}
            line["reference_gpu_kernels"] = {w: r for w, r in refs.items() if r}
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
