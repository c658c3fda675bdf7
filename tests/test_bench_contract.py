"""bench.py's driver contract on the CPU: the reference arm (the CPU oracle port; the one leg that needs no
GPU) prints exactly one JSON line with the keys the driver reads, and the GPU arm refuses to run without a
device instead of falling back to anything."""
import json
import os
import subprocess
import sys

from helpers import ROOT


def _bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), cwd=ROOT, env=e,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)


def test_reference_arm_line(built):
    r = _bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "c1")
    assert r.returncode == 0, r.stderr[-800:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "GStencil/s" and d["unit"] == "GStencil/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 1000          # one step = a sample of about three seconds
    assert d["config"]["workload"].startswith("c1: 2d5pt_star fp64 4096^2")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sub-grid 4096x4096" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GStencil/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_uses_every_host_core_under_torchrun_env(built):
    """torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the CPU legs must not inherit that
    (round 1: the N > 1 reference arm ran on one core and inflated every N > 1 ratio ~15x)."""
    r = _bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "c1", "--gpus", "2",
               env={"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    assert r.returncode == 0, r.stderr[-800:]
    d = json.loads(r.stdout.strip())
    cb = d["cpu_baseline"]
    cores = len(os.sched_getaffinity(0))
    assert cb["host_cores"] == cores and cb["cores"] == cores, cb
    assert "warning" not in cb


def test_reference_arm_other_ranks_stay_silent(built):
    r = _bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "c1", "--gpus", "2",
               env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        return
    r = _bench("--steps", "1")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert r.stdout.strip() == ""
