"""Exploratory timing on the GPU box: device-resident sweeps of the BASELINE configs with
CUDA events, plus the reference's own emitted dr_ kernels (oracle/_ref).  Development aid,
not the contract bench (bench.py)."""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import drstencil_b200 as drs
from drstencil_b200.presets import PRESETS

PEAK = 6553.6


def time_plan(plan, shape, dtype, sweeps=int(os.environ.get('PROBE_SWEEPS', '10')), warm=3):
    A = torch.rand(shape, dtype=dtype, device="cuda")
    B = torch.zeros_like(A)
    bufs = [A, B]
    for s in range(warm):
        plan.sweep(bufs[s & 1], bufs[(s & 1) ^ 1])
    plan.sync_check()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(sweeps):
        plan.sweep(bufs[s & 1], bufs[(s & 1) ^ 1])
    e1.record()
    plan.sync_check()
    return e0.elapsed_time(e1) / sweeps


def report(tag, plan, shape, dtype, ms):
    info = plan.info
    pts = 1
    for n in shape:
        pts *= (n - 2 * info.halo)
    allpts = 1
    for n in shape:
        allpts *= n
    es = 8 if dtype == torch.float64 else 4
    gst = pts * info.step / (ms * 1e-3) / 1e9
    gbs = allpts * 2 * es / (ms * 1e-3) / 1e9
    print("%-28s %8.3f ms  %8.1f GStencil/s  %7.1f GB/s algorithmic (%.1f%% of %.0f)  regs %d smem %d grid %d red %.3f  %s"
          % (tag, ms, gst, gbs, 100 * gbs / PEAK, PEAK, info.regs_per_thread, info.smem_bytes, info.grid_x,
             info.redundancy, plan.knobs), flush=True)
    return dict(tag=tag, ms=ms, gstencil=gst, gbs=gbs, frac=gbs / PEAK, knobs=repr(plan.knobs))


def main():
    which = sys.argv[1:] or ["c1", "c2", "c3", "c4"]
    out = []
    variants = {
        "c1": [dict(sn=128, warps=2, vectors=2, stages=2, rows_per_stage=4)],
        "c2": [dict(step=4, sn=256, vectors=2, stages=2)],
        "c3": [dict(dtype="f32", sn=32, warps=2, rows_per_stage=8, stages=2, min_blocks=4)],
        "c4": [dict(sn=16, rows_3d=4), dict(step=2), dict(step=2, stages=4)],
        "c5": [dict(step=2), dict(step=2, stages=4), dict(step=2, sn=64), dict(step=2, sn=64, stages=4), dict(step=2, sn=256),
               dict(step=2, sn=32), dict(step=2, warps=4, min_blocks=4), dict(step=2, warps=4, min_blocks=4, stages=4)],
    }
    for cfg in which:
        path, _ = PRESETS[cfg]
        for kn in variants[cfg]:
            st = drs.Stencil.from_file(path)
            try:
                plan = drs.Plan(st, drs.Knobs(**kn))
                dtype = torch.float32 if kn.get("dtype") == "f32" else torch.float64
                ms = time_plan(plan, st.shape, dtype)
                out.append(report("%s %s" % (cfg, os.path.basename(path)[3:-4]), plan, st.shape, dtype, ms))
            except Exception as e:
                print(cfg, kn, "FAILED:", str(e)[:300], flush=True)
            torch.cuda.empty_cache()
    # the reference's own emitted kernels on this GPU
    refdir = os.path.join(ROOT, "oracle", "_ref")
    meta = json.load(open(os.path.join(refdir, "cases.json"))) if os.path.exists(os.path.join(refdir, "cases.json")) else {}
    for case, m in sorted(meta.items()):
        if not case.startswith("full_"):
            continue
        lib = ctypes.CDLL(os.path.join(refdir, m["so"]))
        lib.drs_ref_time.restype = ctypes.c_float
        lib.drs_ref_time.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
        for which_k, nm in ((1, "dr_"), (0, "gold_")):
            ms = lib.drs_ref_time(which_k, 6, 2) / 6
            pts = (m["M"] - 2 * 0) * m["N"] * m["L"]
            gst = pts * m["step"] / (ms * 1e-3) / 1e9
            print("REFERENCE %-12s %-5s %8.3f ms  %8.1f GStencil/s (fused step %d)  opts %s"
                  % (case, nm, ms, gst, m["step"], " ".join(m["options"])), flush=True)
            out.append(dict(tag="ref_%s_%s" % (case, nm), ms=ms, gstencil=gst))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
