"""libdrstencil.so loads without a GPU and exports every function include/drstencil.h declares."""
import ctypes
import os
import re

from helpers import ROOT


def test_every_declared_symbol_is_exported(built):
    hdr = open(os.path.join(ROOT, "include", "drstencil.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(drs_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 35
    lib = ctypes.CDLL(os.path.join(ROOT, "drstencil_b200", "libdrstencil.so"))
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


def test_version_and_defaults(built):
    import drstencil_b200 as drs
    assert b"sm_100a" in drs.lib().drs_version()
    k = drs.Knobs().to_c()
    # main.cpp:12-56
    assert (k.step, k.dist, k.streaming, k.bx, k.by, k.sn, k.stream_unroll) == (1, 0, 0, 16, 16, 16, 4)
    assert (k.block_merge_x, k.block_merge_y, k.cyclic_merge_x, k.cyclic_merge_y) == (1, 1, 1, 1)
    assert (k.prefetch, k.merge_forward, k.check) == (0, 5, 0)


def test_product_never_touches_the_oracle():
    """The shipped package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "drstencil_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cuh", ".h", ".cu")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)
