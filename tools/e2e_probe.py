#!/usr/bin/env python
"""Times drs_run_host (host grid in, host grid out) for several block thicknesses of the streamed
path.  usage: python tools/e2e_probe.py [preset] [timesteps] [block ...]   (-1 = plain sequence)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import drstencil_b200 as drs
    from drstencil_b200.presets import PRESETS
    preset = sys.argv[1] if len(sys.argv) > 1 else "c5"
    timesteps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    blocks = [int(x) for x in sys.argv[3:]] or [-1, 0]
    path, kn = PRESETS[preset]
    st = drs.Stencil.from_file(path)
    plan = drs.Plan(st, kn)
    dtype = torch.float32 if kn.dtype == drs.F32 else torch.float64
    h = torch.empty(st.shape, dtype=dtype, pin_memory=True)
    h.fill_(1e-100 if dtype == torch.float64 else 1e-30)
    plan.run_host(h, None, 2)
    for b in blocks:
        plan.set_host_block(b)
        l0 = plan.launch_count
        ms = min(plan.run_host(h, None, timesteps) for _ in range(2))
        print("%s timesteps %d block %5d: %9.2f ms  (%d launches)" % (preset, timesteps, b, ms, plan.launch_count - l0),
              flush=True)


if __name__ == "__main__":
    main()
