"""Head-to-head on the reference's own benchmark set: the eight shipped stencils at the shipped
sizes (8192^2, 512^3) and the reference tuner's fixed --step 2.  Ours = temporal depth 2 (2D) /
composed operator (3D) with default knobs; reference = its emitted dr_ kernel (oracle/_ref,
nvcc sm_100a).  Writes gpurun_out/shipped_vs_reference.{json,md}."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import drstencil_b200 as drs

meta = json.load(open(os.path.join(ROOT, "oracle", "_ref", "cases.json")))
rows = []
for case, m in sorted(meta.items()):
    if not case.startswith("ship_"):
        continue
    name = m["stencil"]
    shape = (m["L"], m["M"], m["N"]) if m["is3d"] else (m["M"], m["N"])
    res = {"stencil": name, "shape": list(shape), "step": m["step"]}
    for label, kn in (("ours_step2", dict(step=2)), ("ours_step1", dict(step=1))):
        st = drs.Stencil.from_file(os.path.join(ROOT, "stc", name + ".stc"))
        plan = drs.Plan(st, drs.Knobs(**kn))
        A = torch.rand(shape, dtype=torch.float64, device="cuda")
        B = torch.zeros_like(A)
        bufs = [A, B]
        for s in range(3):
            plan.sweep(bufs[s & 1], bufs[(s & 1) ^ 1])
        plan.sync_check()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(10):
            plan.sweep(bufs[s & 1], bufs[(s & 1) ^ 1])
        e1.record()
        plan.sync_check()
        ms = e0.elapsed_time(e1) / 10
        H = plan.halo
        pts = 1
        for n in shape:
            pts *= n - 2 * H
        res[label] = {"ms": ms, "gstencil": pts * kn["step"] / (ms * 1e-3) / 1e9, "note": plan.note}
        del A, B, plan
        torch.cuda.empty_cache()
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", m["so"]))
    lib.drs_ref_time.restype = ctypes.c_float
    lib.drs_ref_time.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    ms = lib.drs_ref_time(1, 6, 2) / 6
    H = 2 * (2 if name in ("2d9pt_star", "2d9pt_cross", "2d25pt_box") else 1)
    pts = 1
    for n in shape:
        pts *= n - 2 * H
    res["reference_dr"] = {"ms": ms, "gstencil": pts * 2 / (ms * 1e-3) / 1e9, "options": " ".join(m["options"])}
    rows.append(res)
    print("%-12s ours step2 %8.1f  ours step1 %8.1f  reference dr_ (step 2) %8.1f GStencil/s   x%.2f" %
          (name, res["ours_step2"]["gstencil"], res["ours_step1"]["gstencil"], res["reference_dr"]["gstencil"],
           res["ours_step2"]["gstencil"] / res["reference_dr"]["gstencil"]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "shipped_vs_reference.json"), "w"), indent=1)
with open(os.path.join(ROOT, "gpurun_out", "shipped_vs_reference.md"), "w") as f:
    f.write("# Shipped benchmark set (8192^2 / 512^3, fp64, --step 2 as in the reference's tuning.py) on one B200\n\n")
    f.write("| stencil | ours, step 2 (GStencil/s) | ours, step 1 | reference `dr_` step 2 (its emitted kernel, nvcc sm_100a) | speed-up |\n|---|---|---|---|---|\n")
    for r in rows:
        f.write("| %s | %.1f | %.1f | %.1f | %.2fx |\n" % (r["stencil"], r["ours_step2"]["gstencil"], r["ours_step1"]["gstencil"],
                                                        r["reference_dr"]["gstencil"], r["ours_step2"]["gstencil"] / r["reference_dr"]["gstencil"]))
