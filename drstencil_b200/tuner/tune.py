"""Model-pruned auto-tuner (rebuild of benchmarks/*/tuning.py + starter.sh).

    python -m drstencil_b200.tuner.tune <stc> [--3d] [--step n] [--dtype f32] [--size ...]
                                        [--budget-s 120] [--top 3] [--ncu] [--out results.json]

The reference shuffles its whole Cartesian space and spends an hour running
drstencil -> nvcc -> ncu per candidate (tuning.py:141-160).  Here:
  1. the space is filtered by a resource model (space.filter_config) -- shared memory, register
     window, bytes in flight per SM;
  2. every survivor is specialised in-process (NVRTC, cached) and timed with CUDA events over a
     few sweeps of the real grid; the budget bounds the wall time;
  3. the `--top` best are re-timed and, with --ncu, profiled by Nsight Compute with NAMED metrics
     (tuner/metrics.py) so that each chosen configuration carries its DRAM bytes and GB/s, L2 and
     shared-memory traffic and pipe utilisation -- the evidence BASELINE.json asks for.
Results: JSON (and a `duration.log` in the reference's "elapsed s, best ns" format, tuning.py:104-108).
"""
import argparse
import json
import os
import subprocess
import sys
import time

from .. import F32, Plan, Stencil
from . import metrics as ncu_metrics
from .space import Config, cfg_to_command_line, cfg_to_string, search_space


def time_config(st, cfg: Config, sweeps=6, warm=2):
    import torch
    plan = Plan(st, cfg.knobs())
    dt = torch.float32 if cfg.dtype == "f32" else torch.float64
    A = torch.rand(st.shape, dtype=dt, device="cuda")
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
    ms = e0.elapsed_time(e1) / sweeps
    info = plan.info
    del A, B
    return ms, info


def tune(stc, is3d=None, step=1, dtype="f64", fuse="temporal", size=None, budget_s=120.0, top=3, use_ncu=False,
         peak_gbs=6553.6, log=print, resume_from=None, experimental=False):
    """`resume_from`: a previous result file for the same problem -- configurations already timed
    there are not run again (the reference's tuner always restarts from scratch)."""
    st = Stencil.from_file(stc, is3d)
    if size:
        st.set_size(size)
    radius = max(max(abs(t[0]), abs(t[1]), abs(t[2])) for t in st.terms())
    space = search_space(st.dim, radius, step, dtype, fuse, experimental=experimental)
    log("search space: %d configurations after the resource-model filter" % len(space))
    shape = st.shape
    esize = 4 if dtype == "f32" else 8
    npts = 1
    for n in shape:
        npts *= n
    results = []
    known = {}
    if resume_from and os.path.exists(resume_from):
        try:
            old = json.load(open(resume_from))
            if old.get("stencil") == os.path.basename(stc) and old.get("shape") == list(shape) and \
                    (old.get("step"), old.get("dtype"), old.get("fuse")) == (step, dtype, fuse):
                known = {r["name"]: r for r in old.get("all", [])}
                log("resuming: %d configurations already measured" % len(known))
        except (ValueError, KeyError):
            known = {}
    t0 = time.time()
    best = None
    with open("duration.log", "a") as dl:
        for n, cfg in enumerate(space):
            if cfg_to_string(cfg) in known:
                r = dict(known[cfg_to_string(cfg)])
                r["cfg"] = cfg
                results.append(r)
                continue
            if time.time() - t0 > budget_s:
                log("budget exhausted after %d of %d" % (n, len(space)))
                break
            try:
                ms, info = time_config(st, cfg)
            except Exception as e:   # a configuration the engine refuses is just skipped
                log("%s: skipped (%s)" % (cfg_to_string(cfg), str(e)[:80]))
                continue
            gbs = npts * 2 * esize / (ms * 1e-3) / 1e9
            results.append({"name": cfg_to_string(cfg), "cmd": cfg_to_command_line(cfg), "ms": ms, "gbs": gbs,
                            "frac": gbs / peak_gbs, "regs": info.regs_per_thread, "smem": info.smem_bytes,
                            "grid": info.grid_x, "redundancy": info.redundancy, "cfg": cfg})
            if best is None or ms < best:
                best = ms
                dl.write("%d s, %d\n" % (int(time.time() - t0), int(ms * 1e6)))
            log("%d/%d %s: %.4f ms  %.0f GB/s (%.1f%%)" % (n + 1, len(space), cfg_to_string(cfg), ms, gbs, 100 * gbs / peak_gbs))
    results.sort(key=lambda r: r["ms"])
    winners = results[:top]
    for w in winners:
        ms, _ = time_config(st, w["cfg"], sweeps=20, warm=3)
        w["ms_confirmed"] = ms
        if use_ncu:
            w["ncu"] = profile(stc, st, w["cfg"])
    for r in results:
        r.pop("cfg")
    return {"stencil": os.path.basename(stc), "shape": list(shape), "step": step, "dtype": dtype, "fuse": fuse,
            "peak_gbs": peak_gbs, "tried": len(results), "space": len(space), "seconds": time.time() - t0,
            "winners": winners, "all": results}


def profile(stc, st, cfg: Config):
    """Nsight Compute, named metrics, on tuner/run_one.py for this configuration."""
    size = [str(n) for n in st.shape]
    cmd = ["ncu", "--metrics", ",".join(ncu_metrics.METRICS), "--clock-control", "none", "-k", "regex:dr_", "-s", "2",
           "-c", "3", "--csv", sys.executable, "-m", "drstencil_b200.tuner.run_one", stc] + \
          (["--3d"] if st.dim == 3 else []) + ["--size"] + size + ["--launches", "6", "--"] + cfg_to_command_line(cfg).split()
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    rows = ncu_metrics.parse(r.stdout)
    s = ncu_metrics.summarise(rows)
    if not s:
        return {"error": r.stdout[-400:]}
    return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("stc")
    ap.add_argument("--3d", dest="is3d", action="store_true")
    ap.add_argument("--step", type=int, default=1)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--fuse", default="temporal")
    ap.add_argument("--size", type=int, nargs="+")
    ap.add_argument("--budget-s", type=float, default=120.0)
    ap.add_argument("--top", type=int, default=3)
    ap.add_argument("--ncu", action="store_true")
    ap.add_argument("--out", default="tuning_result.json")
    ap.add_argument("--resume", action="store_true", help="skip configurations already present in --out")
    ap.add_argument("--experimental", action="store_true",
                    help="3D single step: also search six rows per thread, longer chunks and the CTA-shared input ring")
    a = ap.parse_args()
    res = tune(a.stc, a.is3d or None, a.step, a.dtype, a.fuse, a.size, a.budget_s, a.top, a.ncu,
               resume_from=a.out if a.resume else None, experimental=a.experimental)
    json.dump(res, open(a.out, "w"), indent=1)
    for w in res["winners"]:
        print("WINNER %s  %.4f ms  %.0f GB/s (%.1f%% of %.0f)  drstencil%s" %
              (w["name"], w.get("ms_confirmed", w["ms"]), w["gbs"], 100 * w["frac"], res["peak_gbs"], w["cmd"]))


if __name__ == "__main__":
    main()
