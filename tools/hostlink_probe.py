#!/usr/bin/env python
"""Host link under torchrun: H2D / D2H bandwidth of pinned memory per rank, one rank at a time and all
ranks at once (development aid; explains the N > 1 e2e numbers)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
    n = int(gib * 2 ** 30) // 8
    h = torch.empty(n, dtype=torch.float64, pin_memory=True)
    h.fill_(1.0)
    d = torch.empty(n, dtype=torch.float64, device="cuda")

    def timed(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        return gib * 2 ** 30 / (e0.elapsed_time(e1) * 1e-3) / 1e9

    d.copy_(h, non_blocking=True)
    for label, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        for r in range(world):                       # one rank at a time
            dist.barrier()
            if rank == r:
                print("%s rank %d alone      : %6.1f GB/s" % (label, r, timed(fn)), flush=True)
        dist.barrier()
        bw = timed(fn)                               # all ranks at once
        t = torch.tensor([bw], device="cuda")
        dist.all_reduce(t)
        if rank == 0:
            print("%s all %d ranks at once: %6.1f GB/s summed" % (label, world, float(t)), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
