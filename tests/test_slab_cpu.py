"""Host-side logic of the multi-GPU slab path (drstencil_b200/slab.py) with world_size 2 and 3
over gloo on CPU: geometry, the ping-pong schedule and the neighbour halo exchange.  The sweep is
a stand-in supplied by the test (the CPU oracle restricted to the rank's output planes), so what
is checked is that decomposed == undecomposed, bit for bit, including the frozen global ring."""
import os
import socket

import numpy as np
import pytest

from helpers import oracle_terms


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, step, shape, timesteps, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    from drstencil_b200.slab import SlabGeometry, SlabRunner, halo_exchange
    offs, coefs, halo = oracle_terms(name, step)
    L, M, N = shape
    geom = SlabGeometry(L, world, rank, halo)
    a_glob = oracle.rand_array(shape)
    A = np.zeros((geom.local_planes, M, N))
    B = np.zeros_like(A)
    for zl in range(geom.local_planes):
        zg = geom.origin + zl
        if 0 <= zg < L:
            A[zl] = a_glob[zg]
    tA, tB = torch.from_numpy(A), torch.from_numpy(B)

    def sweep(src, dst):
        s, d = src.numpy(), dst.numpy()
        keep = d.copy()
        oracle.sweep(s, d, offs, coefs, halo)            # writes local planes [halo, local_planes - halo)
        # only the rank's output range may change (the global ring stays frozen)
        d[:geom.out_lo] = keep[:geom.out_lo]
        d[geom.out_hi:] = keep[geom.out_hi:]

    runner = SlabRunner(geom, [tA, tB], sweep, lambda dst, s: halo_exchange(dst, geom))
    n = runner.run(timesteps, step)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), tA.numpy()[geom.ghost:geom.ghost + geom.hi - geom.lo])
    np.save(os.path.join(out_dir, "n%d.npy" % rank), np.array([n]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,name,step,shape,timesteps", [
    (2, "3d7pt_star", 1, (12, 10, 12), 4),
    (3, "3d7pt_star", 2, (20, 9, 10), 8),
    (2, "3d9pt_cross", 1, (9, 8, 8), 6),
])
def test_decomposed_equals_global(built, tmp_path, world, name, step, shape, timesteps):
    import torch.multiprocessing as mp
    from oracle import oracle
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, step, shape, timesteps, str(tmp_path)))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = np.concatenate([np.load(tmp_path / ("rank%d.npy" % r)) for r in range(world)], axis=0)
    offs, coefs, halo = oracle_terms(name, step)
    A = oracle.rand_array(shape)
    B = np.zeros(shape)
    n = oracle.run(A, B, offs, coefs, halo, timesteps, step)
    assert int(np.load(tmp_path / "n0.npy")[0]) == n
    assert np.array_equal(got, A)


def test_geometry():
    from drstencil_b200.slab import SlabGeometry, split
    assert split(10, 3) == [(0, 3), (3, 6), (6, 10)]
    g = SlabGeometry(1536, 8, 0, 1)
    assert (g.lo, g.hi, g.local_planes, g.out_lo, g.out_hi, g.lower, g.upper) == (0, 192, 194, 2, 193, None, 1)
    g = SlabGeometry(1536, 8, 7, 1)
    assert (g.lo, g.hi, g.out_lo, g.out_hi, g.upper) == (1344, 1536, 1, 192, None)
    g = SlabGeometry(1536, 8, 3, 2)
    assert g.send_lower() == (2, 4) and g.recv_lower() == (0, 2)
    assert g.send_upper() == (192, 194) and g.recv_upper() == (194, 196)
    with pytest.raises(ValueError):
        SlabGeometry(8, 8, 0, 2)
