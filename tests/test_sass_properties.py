"""Static properties of the compiled sm_100a kernels, read from the SASS of the cubins NVRTC produced
(cuobjdump, no GPU needed): the hot loops are fed by the TMA unit through mbarriers, never by per-thread
global loads; outputs leave as 128-bit stores; the bit-exact (depth 1) kernels contain nothing but the
explicit mul/fma chain -- no floating-point add the compiler could have re-associated."""
import collections
import os
import re
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")


def _mnemonics(preset):
    import drstencil_b200 as drs
    from drstencil_b200.presets import PRESETS
    path, kn = PRESETS[preset]
    plan = drs.Plan(drs.Stencil.from_file(path), kn)
    cubin = os.path.join(os.path.dirname(drs.__file__), "_jitcache", plan.cache_key + ".cubin")
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", plan.info.kernel_name, cubin], stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True).stdout
    ops = collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(1)] += 1
    assert sum(ops.values()) > 200, sass[:400]
    return ops, plan


def _count(ops, prefix):
    return sum(n for k, n in ops.items() if k.startswith(prefix))


@pytest.mark.parametrize("preset", ["c1", "c2", "c3", "c4", "c5", "c4t2", "c1t2"])
def test_tma_mbarrier_pipeline_and_vector_stores(built, preset):
    ops, plan = _mnemonics(preset)
    dim = plan.info.dim
    assert _count(ops, "UTMALDG.%dD" % dim) >= 2                  # cp.async.bulk.tensor: prologue + steady state
    assert _count(ops, "SYNCS.ARRIVE.TRANS64") >= 2               # mbarrier.arrive.expect_tx
    assert _count(ops, "SYNCS.PHASECHK.TRANS64.TRYWAIT") >= 2     # mbarrier.try_wait.parity
    assert _count(ops, "STG.E.128") >= 1                          # one 128-bit store per interior vector
    # the only per-thread global loads are the volatile reads of the watchdog flag (mbar_wait) and, in the 3D
    # kernels, the slab protocol's words (drs_common.cuh: ld.acquire.sys of the step flags, the sequence base):
    # a handful of scalar loads outside the plane loop, never grid data
    ldg = {k: n for k, n in ops.items() if k.startswith("LDG")}
    allowed = {"LDG.E.STRONG.SYS"} | ({"LDG.E.64", "LDG.E.64.STRONG.SYS"} if dim == 3 else set())
    assert set(ldg) <= allowed and sum(ldg.values()) <= (16 if dim == 3 else 8), ldg
    assert _count(ops, "LDL") == 0 and _count(ops, "STL") == 0    # no local-memory traffic (spills)


@pytest.mark.parametrize("preset,mul,fma,add", [
    ("c1", "DMUL", "DFMA", "DADD"), ("c4", "DMUL", "DFMA", "DADD"), ("c5", "DMUL", "DFMA", "DADD"),
    ("c3", "FMUL", "FFMA", "FADD"),
])
def test_bit_exact_kernels_hold_only_the_ordered_chain(built, preset, mul, fma, add):
    """Depth-1 kernels: every output is mul(t2), fma(t1), fma(t3) ...: P - 1 fused multiply-adds
    per multiply, and no add at all."""
    ops, plan = _mnemonics(preset)
    npts = plan.info.npoints
    n_mul, n_fma = _count(ops, mul), _count(ops, fma)
    assert _count(ops, add) == 0
    assert n_mul > 0 and n_fma == (npts - 1) * n_mul, (n_mul, n_fma, npts)


@pytest.mark.parametrize("preset", ["c2", "c1t2", "c4t2"])
def test_temporal_kernels_exchange_neighbours_by_shuffle(built, preset):
    ops, _ = _mnemonics(preset)
    assert _count(ops, "SHFL.UP") >= 1 and _count(ops, "SHFL.DOWN") >= 1
