"""Peer stores of the slab decomposition on ONE GPU: the ranks of a 2- / 3-way decomposition are emulated as plans of
one process whose "neighbour" arrays live on the same device.  Every sweep of every rank is the product kernel with
`drs_plan_set_slab` + `drs_plan_set_peers` (boundary planes stored a second time into the neighbours' ghost planes);
the ranks' sweeps of one step run back to back on one stream, which is all the ordering the exchange needs.  The
owned planes must equal the undecomposed run of the same grid bit for bit (same kernel, same arithmetic) -- the
property bench.py checks across real GPUs at N > 1 (`parity.slab_vs_single`), here for the driver's one-GPU suite."""
import numpy as np
import pytest

from helpers import stc_path

pytestmark = pytest.mark.gpu


def _run_emulated(name, shape, world, sweeps, np_dtype=np.float64, **kn):
    import torch
    import drstencil_b200 as drs
    from drstencil_b200.slab import SlabGeometry
    from oracle import oracle
    L, M, N = shape
    knobs = drs.Knobs(**kn)
    whole = drs.Plan(drs.Stencil.from_file(stc_path(name)).set_size(shape), knobs)
    ghost = whole.halo
    a0 = oracle.rand_array(shape, np_dtype)
    A = torch.from_numpy(a0).cuda()
    B = torch.zeros_like(A)
    bufs = [A, B]
    for s in range(sweeps):
        whole.sweep(bufs[s & 1], bufs[(s & 1) ^ 1])
    whole.sync_check()
    want = bufs[sweeps & 1].cpu().numpy()

    geoms = [SlabGeometry(L, world, r, ghost) for r in range(world)]
    plans, arrays = [], []
    for g in geoms:
        st = drs.Stencil.from_file(stc_path(name)).set_size((g.local_planes, M, N))
        p = drs.Plan(st, knobs)
        p.set_slab(L, g.lo, g.hi)
        la = np.zeros((g.local_planes, M, N), np_dtype)
        for zl in range(g.local_planes):
            zg = g.origin + zl
            if 0 <= zg < L:
                la[zl] = a0[zg]
        plans.append(p)
        arrays.append([torch.from_numpy(la).cuda(), torch.zeros((g.local_planes, M, N), dtype=A.dtype, device="cuda")])
    for r, (g, p) in enumerate(zip(geoms, plans)):
        lower = arrays[r - 1] if g.lower is not None else [0, 0]
        upper = arrays[r + 1] if g.upper is not None else [0, 0]
        p.set_peers(arrays[r], lower, upper, g.bounds[g.lower][0] if g.lower is not None else 0,
                    g.bounds[g.upper][0] if g.upper is not None else 0)
    for s in range(sweeps):
        for r, p in enumerate(plans):
            p.sweep(arrays[r][s & 1], arrays[r][(s & 1) ^ 1])
    for p in plans:
        p.sync_check()
    for r, g in enumerate(geoms):
        got = arrays[r][sweeps & 1][g.ghost:g.ghost + (g.hi - g.lo)].cpu().numpy()
        assert np.array_equal(got, want[g.lo:g.hi]), "rank %d of %d: owned planes differ from the undecomposed run" % (r, world)
    return plans


@pytest.mark.parametrize("name,world,shape,kn", [
    ("3d7pt_star", 2, (40, 48, 72), dict()),
    ("3d7pt_star", 3, (45, 37, 130), dict()),
    ("3d9pt_cross", 2, (24, 40, 64), dict()),
    # the c4 / c5 presets' kernel: four warps share one input ring
    ("3d7pt_star", 3, (50, 64, 128), dict(bx=32, by=4, sn=16, share_x=2, share_y=2, rows_3d=8)),
    # chunks shorter than the ghost depth, thin slabs with one chunk touching both faces
    ("3d9pt_cross", 3, (15, 40, 64), dict(sn=3)),
])
def test_single_step_peer_stores(built, name, world, shape, kn):
    _run_emulated(name, shape, world, 4, **kn)


@pytest.mark.parametrize("name,step,world,shape,kn", [
    ("3d7pt_star", 2, 2, (40, 48, 72), dict()),
    ("3d7pt_star", 2, 3, (47, 70, 130), dict()),
    ("3d7pt_star", 3, 2, (40, 48, 72), dict()),          # Halo 3: partial vectors at the grid edge in x
    ("3d9pt_cross", 2, 2, (36, 48, 72), dict()),
    ("3d7pt_star", 2, 3, (36, 48, 72), dict(sn=5)),
])
def test_fused_temporal_peer_stores(built, name, step, world, shape, kn):
    """drs_sweep3d_t.cuh pushes n * r boundary planes per side."""
    plans = _run_emulated(name, shape, world, 4, step=step, **kn)
    assert all("drs_sweep3d_t.cuh" in p.source for p in plans)


def test_fused_temporal_peer_stores_fp32(built):
    _run_emulated("3d7pt_star", (40, 48, 136), 2, 2, np_dtype=np.float32, step=2, dtype="f32")
