#!/usr/bin/env python
"""Where does the per-sweep cost of the slab protocol go?  Under torchrun (N >= 2) times c5 slabs
with the pieces of the halo protocol switched off one at a time (timing only -- the ablated
variants compute garbage at the slab faces).  Development aid.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/slab_probe.py [planes_per_rank]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from drstencil_b200 import slab as S
    from drstencil_b200.presets import PRESETS
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    path, kn = PRESETS["c5"]
    per = int(sys.argv[1]) if len(sys.argv) > 1 else 1536 // world
    shape = (per * world, 1536, 1536)
    sweeps = 100
    for variant in ("full", "noflags", "nopush", "neither", "full"):
        sl = S.GpuSlab(path, kn, rank, world, halo="p2p", global_shape=shape)
        if variant in ("nopush", "neither"):
            sl.plan.set_peers([r.ptr for r in sl._raw], [0, 0], [0, 0], 0, 0)
        if variant in ("noflags", "neither"):
            sl._before = lambda s: None
            sl._after = lambda dst, s: None
            sl.runner.before, sl.runner.after = sl._before, sl._after
        sl.bufs[0].fill_(1e-200)
        sl.run(10)
        sl.plan.sync_check()
        best = None
        for _ in range(3):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sl.run(sweeps)
            e1.record()
            sl.plan.sync_check()
            t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t) if best is None else min(best, float(t))
        if rank == 0:
            print("%-8s %d planes/rank: %8.2f us per sweep" % (variant, per, best / sweeps * 1e3), flush=True)
        sl.close()
        for r in sl._raw:
            r.free()
        del sl
        torch.cuda.empty_cache()
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
