#!/usr/bin/env python
"""Host link under torchrun: H2D / D2H bandwidth of pinned memory per rank -- one rank at a time, all ranks at
once, one and two copy streams per direction, both directions at once (development aid: explains, and bounds,
the N > 1 e2e numbers).  Prints one table and writes gpurun_out/hostlink_n<world>.json."""
import json
import os
import sys

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
    h2 = torch.empty(n, dtype=torch.float64, pin_memory=True)
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    d2 = torch.empty(n, dtype=torch.float64, device="cuda")
    streams = [torch.cuda.Stream() for _ in range(4)]
    half = n // 2

    def h2d(k):
        if k == 1:
            d.copy_(h, non_blocking=True)
        else:
            for i, s in enumerate(streams[:2]):
                with torch.cuda.stream(s):
                    d[i * half:(i + 1) * half].copy_(h[i * half:(i + 1) * half], non_blocking=True)

    def d2h(k):
        if k == 1:
            h2.copy_(d2, non_blocking=True)
        else:
            for i, s in enumerate(streams[2:]):
                with torch.cuda.stream(s):
                    h2[i * half:(i + 1) * half].copy_(d2[i * half:(i + 1) * half], non_blocking=True)

    def both(k):
        with torch.cuda.stream(streams[0]):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(streams[2]):
            h2.copy_(d2, non_blocking=True)

    def timed(fn, k, nbytes):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        e0.record()
        for s in streams:
            s.wait_stream(cur)
        fn(k)
        for s in streams:
            cur.wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        return nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9

    one = gib * 2 ** 30
    h2d(1)
    d2h(1)
    torch.cuda.synchronize()
    out = {"world": world, "gib_per_rank": gib, "rows": []}
    for label, fn, k, nbytes in (("H2D, 1 stream", h2d, 1, one), ("H2D, 2 streams", h2d, 2, one), ("D2H, 1 stream", d2h, 1, one),
                                 ("D2H, 2 streams", d2h, 2, one), ("H2D + D2H at once", both, 1, 2 * one)):
        alone = []
        for r in range(world):                       # one rank at a time
            dist.barrier()
            bw = timed(fn, k, nbytes) if rank == r else 0.0
            t = torch.tensor([bw], device="cuda")
            dist.all_reduce(t)
            alone.append(float(t))
        dist.barrier()
        bw = timed(fn, k, nbytes)                    # all ranks at once
        t = torch.tensor([bw], device="cuda")
        dist.all_reduce(t)
        tmin = torch.tensor([bw], device="cuda")
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        row = {"what": label, "alone_gbs_per_rank": alone, "all_at_once_gbs_sum": float(t), "all_at_once_gbs_min_rank": float(tmin)}
        out["rows"].append(row)
        if rank == 0:
            print("%-20s alone: %s GB/s | all %d ranks at once: %6.1f GB/s summed, slowest rank %5.1f"
                  % (label, " ".join("%5.1f" % a for a in alone), world, float(t), float(tmin)), flush=True)
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "hostlink_n%d.json" % world), "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
