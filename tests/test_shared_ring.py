"""The CTA-shared input ring of the single-step 3D sweep (drs_sweep3d_cta.cuh; engine override share_x / share_y):
the c5 preset since round 2 (2 x 2 warps per ring, eight rows per thread: 386.6 vs 376.8 GStencil/s sustained on
1536^3 for the private rings).  Compiled and resource-checked here on the CPU; bit-exact parity on the GPU."""
import os

import numpy as np
import pytest

from helpers import oracle_run, stc_path


def _plan(name, shape, **kn):
    import drstencil_b200 as drs
    st = drs.Stencil.from_file(stc_path(name)).set_size(shape)
    return drs.Plan(st, drs.Knobs(**kn))


def test_shared_ring_specialisation_compiles_and_reports_its_geometry(built):
    import drstencil_b200 as drs
    shape = (64, 1536, 1536)
    base = _plan("3d7pt_star", shape, sn=64, rows_3d=6, warps=4)
    assert "drs_sweep3d.cuh" in base.source and "drs_sweep3d_cta.cuh" not in base.source
    for sx, sy in [(2, 2), (2, 1), (1, 2), (3, 2)]:
        plan = _plan("3d7pt_star", shape, sn=64, rows_3d=6, share_x=sx, share_y=sy)
        info = plan.info
        assert "drs_sweep3d_cta.cuh" in plan.source
        assert "#define DRS_SX %d\n#define DRS_SY %d\n" % (sx, sy) in plan.source
        assert info.warps_per_cta == sx * sy
        # one ring per CTA: 4 stages of (sx*64 + 2) x (sy*6 + 2) doubles + full/empty barriers
        stage = ((sx * 64 + 4) * (sy * 6 + 2) * 8 + 127) // 128 * 128
        assert info.smem_bytes == 4 * stage + 2 * 4 * 8
        # CTA grid: ceil over the warp tiles per axis
        nxs, nys, nzs = 1536 // 64, -(-(1536 - 2) // 6), 1
        assert info.grid_x == -(-nxs // sx) * -(-nys // sy) * nzs
        log = open(os.path.join(os.path.dirname(drs.__file__), "_jitcache", plan.cache_key + ".log")).read()
        assert "0 bytes spill stores, 0 bytes spill loads" in log
    # four warps sharing one ring need less shared memory than four private rings
    assert _plan("3d7pt_star", shape, sn=64, rows_3d=6, share_x=2, share_y=2).info.smem_bytes < base.info.smem_bytes


def test_shared_ring_argument_errors(built):
    import drstencil_b200 as drs
    with pytest.raises(drs.DrsError, match="256-element TMA box"):
        _plan("3d7pt_star", (64, 512, 512), dtype="f32", share_x=2, share_y=2)     # 2 * 128 + 8 columns
    # 2D plans and temporal 3D plans ignore the override
    assert "drs_sweep2d.cuh" in _plan("2d5pt_star", (256, 256), share_x=2, share_y=2).source
    assert "drs_sweep3d_t.cuh" in _plan("3d7pt_star", (64, 128, 128), step=2, share_x=2, share_y=2).source


@pytest.mark.gpu
@pytest.mark.parametrize("name,shape,kn", [
    ("3d7pt_star", (40, 48, 264), dict(share_x=2, share_y=2)),
    ("3d7pt_star", (19, 21, 66), dict(share_x=2, share_y=2, sn=5)),          # warps beyond both edges
    ("3d7pt_star", (30, 50, 130), dict(share_x=3, share_y=1, rows_3d=6)),
    ("3d7pt_star", (30, 50, 130), dict(share_x=1, share_y=3, rows_3d=4)),
    ("3d9pt_cross", (24, 40, 200), dict(share_x=2, share_y=2)),
    ("3d7pt_star", (24, 40, 264), dict(share_x=1, share_y=2, dtype="f32")),
    ("3d7pt_star", (70, 50, 264), dict(share_x=2, share_y=2, rows_3d=8, sn=64)),   # the c5 preset's shape
    ("3d7pt_star", (40, 30, 200), dict(share_x=2, share_y=2, rows_3d=12, sn=16)),
])
def test_shared_ring_bit_exact(built, name, shape, kn):
    import torch
    from oracle import oracle
    dtype = np.float32 if kn.get("dtype") == "f32" else np.float64
    plan = _plan(name, shape, **kn)
    assert "drs_sweep3d_cta.cuh" in plan.source
    a0 = oracle.rand_array(shape).astype(dtype)
    A, B = torch.from_numpy(a0).cuda(), torch.full(shape, -3.0, dtype=torch.from_numpy(a0).dtype, device="cuda")
    plan.sweep(A, B)
    plan.sweep(B, A)
    plan.sync_check()
    from helpers import oracle_terms
    offs, coefs, halo = oracle_terms(name, 1)
    refA, refB = a0.copy(), np.full(shape, -3.0, dtype)
    oracle.sweep(refA, refB, offs, coefs, halo)
    oracle.sweep(refB, refA, offs, coefs, halo)
    assert np.array_equal(B.cpu().numpy(), refB) and np.array_equal(A.cpu().numpy(), refA)
