"""Run under torchrun (one process per GPU): the slab-decomposed sweep against the undecomposed
single-GPU run of the same global grid, both exchange paths.  Prints SLAB_CHECK_OK on success."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import drstencil_b200 as drs
from drstencil_b200.slab import GpuSlab


def host_slabs(rank, world):
    """GpuSlab.run_host -- the streamed host-buffer run of a slab-decomposed grid (drs_run_host_slab) -- against
    the undecomposed single-GPU run, two calls in a row (flag values carry over), then a device-resident
    run() on the same object."""
    ok = True
    for name, shape, kn, timesteps, block in [
        ("3d7pt_star", (24 * world, 40, 128), dict(), 8, 6),
        ("3d7pt_star", (30 * world + 1, 33, 66), dict(sn=5, rows_3d=4), 20, 4),
        ("3d7pt_star", (32 * world, 44, 130), dict(step=2, sn=9), 8, 8),
        ("3d7pt_star", (96 * world, 256, 256), dict(sn=16), 12, 0),          # engine-chosen blocks
    ]:
        path = os.path.join(ROOT, "stc", name + ".stc")
        L, M, N = shape
        g = torch.Generator(device="cuda")

        def plane(zg):
            g.manual_seed(99 + zg)
            return torch.rand((M, N), dtype=torch.float64, device="cuda", generator=g)

        st = drs.Stencil.from_file(path).set_size(shape)
        plan = drs.Plan(st, drs.Knobs(**kn))
        A = torch.stack([plane(z) for z in range(L)])
        B = torch.zeros_like(A)
        plan.run(A, B, timesteps)
        plan.sync_check()
        slab = GpuSlab(path, drs.Knobs(**kn), rank, world, halo="p2p", global_shape=shape)
        if block:
            slab.plan.set_host_block(block)
        lo, hi = slab.geom.lo, slab.geom.hi
        same = True
        for call in range(2):
            h = torch.stack([plane(z) for z in range(lo, hi)]).cpu().pin_memory()
            slab.run_host(h, timesteps)
            same = same and bool(torch.equal(h.cuda(), A[lo:hi]))
        # the same object keeps working device-resident afterwards (one flag numbering for both calls)
        slab.fill(plane)
        slab.run(timesteps)
        slab.plan.sync_check()
        same = same and bool(torch.equal(slab.owned(0), A[lo:hi]))
        t = torch.tensor([1 if same else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("slab_check run_host %s %s %s block=%d world=%d -> %s" % (name, shape, kn, block, world,
                                                                             "bit-exact" if int(t) else "MISMATCH"), flush=True)
        ok = ok and bool(int(t))
        slab.close()
        dist.barrier()
    return ok


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for name, shape, kn, timesteps in [
        ("3d7pt_star", (64, 40, 128), dict(), 8),
        ("3d7pt_star", (37, 33, 66), dict(sn=5, rows_3d=4), 4),
        ("3d9pt_cross", (48, 24, 64), dict(), 4),
        ("3d7pt_star", (40, 40, 64), dict(step=2, fuse="algebraic"), 8),   # composed operator, ghost = 2
        ("3d7pt_star", (56, 44, 130), dict(step=2, sn=9), 8),              # fused temporal kernel, ghost = 2
        ("3d9pt_cross", (60, 40, 64), dict(step=2), 8),
        ("3d7pt_star", (8 * world + 3, 96, 264), dict(share_x=2, share_y=2, sn=3), 8),   # CTA-shared ring, thin slabs: chunks on both faces
        ("3d7pt_star", (16 * world, 40, 128), dict(sn=1), 8),              # one plane per chunk
        ("3d7pt_star", (24 * world, 40, 128), dict(step=2, sn=1), 8),      # ghost = 2 > chunk: two chunks per face
        ("3d7pt_star", (48 * world, 384, 384), dict(sn=8, rows_3d=6, share_x=2, share_y=2), 200),   # many sweeps, several waves
        ("3d7pt_star", (48 * world, 384, 384), dict(sn=8, rows_3d=4), 200),
        ("3d7pt_star", (48 * world, 384, 384), dict(step=2, sn=16), 200),
    ]:
        path = os.path.join(ROOT, "stc", name + ".stc")
        L, M, N = shape
        g = torch.Generator(device="cuda")

        def plane(zg):
            g.manual_seed(99 + zg)
            return torch.rand((M, N), dtype=torch.float64, device="cuda", generator=g)

        # undecomposed reference on this GPU
        st = drs.Stencil.from_file(path).set_size(shape)
        plan = drs.Plan(st, drs.Knobs(**kn))
        A = torch.stack([plane(z) for z in range(L)])
        B = torch.zeros_like(A)
        plan.run(A, B, timesteps)
        plan.sync_check()
        for mode in ("p2p", "p2p-flags", "nccl"):
            if mode != "p2p" and timesteps > 50:
                continue
            slab = GpuSlab(path, drs.Knobs(**kn), rank, world, halo=mode, global_shape=shape)
            slab.fill(plane)
            assert timesteps % (4 * kn.get("step", 1)) == 0
            slab.run(timesteps // 2)          # two calls: the step flags carry over between them
            slab.run(timesteps // 2)
            slab.plan.sync_check()
            torch.cuda.synchronize()
            dist.barrier()
            mine = slab.owned(0)
            ref = A[slab.geom.lo:slab.geom.hi]
            same = bool(torch.equal(mine, ref))
            t = torch.tensor([1 if same else 0], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if rank == 0:
                print("slab_check %s %s %s mode=%s world=%d -> %s" % (name, shape, kn, mode, world,
                                                                       "bit-exact" if int(t) else "MISMATCH"), flush=True)
            ok = ok and bool(int(t))
            slab.close()
            dist.barrier()
    if os.environ.get("DRS_SLAB_CHECK_HOST", "1") == "1":
        ok = host_slabs(rank, world) and ok
    if rank == 0:
        print("SLAB_CHECK_OK" if ok else "SLAB_CHECK_FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
