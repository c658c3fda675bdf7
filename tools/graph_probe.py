#!/usr/bin/env python
"""c1-sized schedules: stream launches vs one CUDA graph of the same launches (development aid)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import drstencil_b200 as drs
    from drstencil_b200.presets import PRESETS
    preset = sys.argv[1] if len(sys.argv) > 1 else "c1"
    ts = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    path, kn = PRESETS[preset]
    st = drs.Stencil.from_file(path)
    plan = drs.Plan(st, kn)
    dtype = torch.float32 if kn.dtype == drs.F32 else torch.float64
    A = torch.rand(st.shape, dtype=dtype, device="cuda") * 1e-100
    B = torch.zeros_like(A)
    plan.run(A, B, ts)
    plan.sync_check()

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    t_stream = timed(lambda: plan.run(A, B, ts))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        plan.run(A, B, ts)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        plan.run(A, B, ts)
    t_graph = timed(g.replay)
    n = drs.sweep_count(ts, kn.step)
    print("%s %d timesteps (%d launches): stream %.2f us/launch, graph %.2f us/launch" %
          (preset, ts, n, t_stream / n * 1e3, t_graph / n * 1e3))


if __name__ == "__main__":
    main()
