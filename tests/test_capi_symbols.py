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


def test_one_pinned_compiler_whatever_was_imported_first(built):
    """Round 1: the library dlopen()ed "libnvrtc.so.12", which resolved to torch's copy (12.8) or the toolkit's (12.9)
    depending on import order -- two compilers, two sets of cubins.  Now one absolute path is compiled in."""
    import subprocess
    import sys
    import drstencil_b200 as drs
    with_torch = drs.lib().drs_compiler().decode()
    code = ("import ctypes, sys; assert 'torch' not in sys.modules; L = ctypes.CDLL(%r); "
            "L.drs_compiler.restype = ctypes.c_char_p; print(L.drs_compiler().decode())"
            % os.path.join(ROOT, "drstencil_b200", "libdrstencil.so"))
    alone = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, text=True, check=True).stdout.strip()
    assert alone == with_torch and alone.startswith("NVRTC ") and "(/" in alone
