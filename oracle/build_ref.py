#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- builds the *real* reference into oracle/_ref/ (git-ignored).

Run in the authoring container, where /root/reference exists:

  1. oracle/_ref/drstencil_ref          g++ -O3 -std=c++17 /root/reference/main.cpp  (the
                                        reference's own Makefile line, Makefile:7)
  2. for every case in CASES: a size-edited .stc (the reference bakes sizes into the emitted
     program, codegen_2d.hpp:87-90), the reference-emitted `--check` program
     oracle/_ref/cases/<case>/<name>.cu, and oracle/_ref/libref_<case>.so = ref_wrap.cu around it,
     compiled with the reference's nvcc flags (benchmarks/2d5pt_star/compile_run.sh:4)
     retargeted from sm_80 to sm_100a.
  3. oracle/_ref/cases.json             what was built, for the tests / bench to enumerate.

Nothing from /root/reference is copied into the repository: sources are compiled where they lie
(common.hpp through -I), only binaries and the generator's *output* land in oracle/_ref/.
The GPU box has no /root/reference; it uses the prebuilt files shipped with the snapshot.
"""
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"

# name -> (stencil file stem, is3d, (L, M, N), iterations, step, extra generator options)
# Small cases: parity (bit-exactness of the oracle / product vs the reference's gold kernel).
# "full_*" cases: the BASELINE.json sizes, used to time the reference's own dr_ kernel on B200.
CASES = {
    "p2d5s":   ("2d5pt_star", False, (1, 200, 264), 4, 1, ["--streaming", "--bx", "64", "--sn", "16"]),
    "p2d5s2":  ("2d5pt_star", False, (1, 200, 264), 4, 2, ["--streaming", "--bx", "64", "--sn", "16"]),
    "p2d5k1":  ("2d5pt_star", False, (1, 130, 150), 4, 1, []),
    "p2d5x":   ("2d5pt_cross", False, (1, 200, 264), 4, 1, ["--dist", "2"]),
    "p2d9s":   ("2d9pt_star", False, (1, 200, 264), 4, 1, ["--streaming", "--bx", "64", "--sn", "16"]),
    "p2d9b":   ("2d9pt_box", False, (1, 200, 264), 4, 1, ["--streaming", "--bx", "64", "--sn", "16"]),
    "p2d9b2":  ("2d9pt_box", False, (1, 200, 264), 4, 2, ["--streaming", "--bx", "64", "--sn", "16"]),
    "p2d9b4":  ("2d9pt_box", False, (1, 200, 264), 8, 4, ["--streaming", "--bx", "64", "--sn", "16"]),
    "p2d9x":   ("2d9pt_cross", False, (1, 200, 264), 4, 1, ["--dist", "2", "--streaming", "--bx", "64", "--sn", "16"]),
    "p2d25b":  ("2d25pt_box", False, (1, 200, 264), 4, 1, ["--streaming", "--bx", "64", "--sn", "16"]),
    "p3d7s":   ("3d7pt_star", True, (40, 48, 72), 4, 1, []),
    "p3d7s2":  ("3d7pt_star", True, (40, 48, 72), 4, 2, []),
    "p3d9x":   ("3d9pt_cross", True, (40, 48, 72), 4, 1, ["--dist", "2"]),
    "full_c1": ("2d5pt_star", False, (1, 4096, 4096), 10, 1,
                ["--streaming", "--bx", "128", "--sn", "64", "--cyclic-merge-x", "2", "--prefetch"]),
    "full_c2": ("2d9pt_box", False, (1, 16384, 16384), 8, 4,
                ["--streaming", "--bx", "128", "--sn", "64", "--cyclic-merge-x", "2", "--prefetch"]),
    "full_c2s1": ("2d9pt_box", False, (1, 16384, 16384), 8, 1,
                  ["--streaming", "--bx", "128", "--sn", "64", "--cyclic-merge-x", "2", "--prefetch"]),
    "full_c3f64": ("2d25pt_box", False, (1, 16384, 16384), 4, 1,
                   ["--streaming", "--bx", "128", "--sn", "64", "--cyclic-merge-x", "2", "--prefetch"]),
    "full_c4": ("3d7pt_star", True, (768, 768, 768), 4, 1, ["--bx", "32", "--by", "8", "--sn", "32"]),
}
# the eight shipped descriptions at the reference's own sizes (8192^2 / 512^3) and its tuner's fixed
# --step 2 (benchmarks/*/tuning.py: `range(2, 3)`), typical streaming options: head-to-head table
_S2D = ["--streaming", "--bx", "128", "--sn", "64", "--cyclic-merge-x", "2", "--prefetch"]
for _n in ("2d5pt_star", "2d5pt_cross", "2d9pt_star", "2d9pt_box", "2d9pt_cross", "2d25pt_box"):
    CASES["ship_" + _n] = (_n, False, (1, 8192, 8192), 4, 2, _S2D)
for _n in ("3d7pt_star", "3d9pt_cross"):
    CASES["ship_" + _n] = (_n, True, (512, 512, 512), 4, 2, ["--bx", "32", "--by", "8", "--sn", "32"])


# the winners of the reference's own search space on B200 (oracle/tune_ref.py; table: profiles/r02_ref_tune.json):
# bench.py times these as `reference_gpu_kernels` instead of round 1's hand-picked options
_BEST = os.path.join(HERE, "ref_best.json")
_FULL = {"c1": ("2d5pt_star", False, (1, 4096, 4096), 10, 1), "c2": ("2d9pt_box", False, (1, 16384, 16384), 8, 4),
         "c4": ("3d7pt_star", True, (768, 768, 768), 4, 1)}
if os.path.exists(_BEST):
    for _wl, _b in json.load(open(_BEST)).items():
        if _wl in _FULL:
            _stem, _is3d, _dims, _iters, _step = _FULL[_wl]
            CASES["best_" + _wl] = (_stem, _is3d, _dims, _iters, _step, list(_b["options"]))


def sh(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    return r.returncode, r.stdout


def opt(opts, name, default):
    return int(opts[opts.index(name) + 1]) if name in opts else default


def generator():
    """oracle/_ref/drstencil_ref: the reference generator, built with its own Makefile line (Makefile:7)."""
    os.makedirs(os.path.join(OUT, "cases"), exist_ok=True)
    gen = os.path.join(OUT, "drstencil_ref")
    if not os.path.exists(gen):
        rc, out = sh(["g++", "-O3", "-std=c++17", "-w", "-o", gen, os.path.join(REF, "main.cpp")])
        if rc != 0:
            raise RuntimeError("reference generator failed to build:\n" + out)
    return gen


def build_one(case, stem, is3d, dims, iters, step, opts, cdir, so):
    """One reference-emitted `--check` program (size-edited .stc -> drstencil_ref -> nvcc with the reference's
    flags, sm_100a) wrapped by ref_wrap.cu into `so`.  Returns its metadata."""
    gen = generator()
    L, M, N = dims
    os.makedirs(cdir, exist_ok=True)
    # coefficient table taken from the shipped .stc, sizes edited in
    src = os.path.join(REF, "benchmarks", stem, stem + ".stc")
    toks = open(src).read().split()
    body = toks[toks.index("stencil") + 1:]
    w = 4 if is3d else 3
    with open(os.path.join(cdir, stem + ".stc"), "w") as f:
        if is3d:
            f.write("L %d\n" % L)
        f.write("M %d\nN %d\n\niterations %d\n\nstencil\n" % (M, N, iters))
        for q in range(0, len(body), w):
            f.write(" ".join(body[q:q + w]) + "\n")
    args = [gen] + (["--3d"] if is3d else []) + ["--step", str(step)] + opts + \
           ["--check", "-o", stem + ".cu", stem + ".stc"]
    rc, out = sh(args, cwd=cdir)
    cu = os.path.join(cdir, stem + ".cu")
    if rc != 0 or not os.path.exists(cu):
        raise RuntimeError("reference generator failed on %s (rc %d): %s" % (case, rc, out))
    streaming = 1 if "--streaming" in opts else 0
    mx = max(opt(opts, "--block-merge-x", 1), opt(opts, "--cyclic-merge-x", 1))
    my = max(opt(opts, "--block-merge-y", 1), opt(opts, "--cyclic-merge-y", 1))
    cmd = [NVCC, "-maxrregcount=128", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++11",
           "--use_fast_math", "-Xptxas", "-dlcm=cg", "-w", "-shared", "-Xcompiler", "-fPIC",
           "-Xlinker", "-Bsymbolic", "-I", REF, "-I", cdir,
           '-DDRS_REF_CU="%s"' % cu, "-DDRS_REF_NAME=" + stem, "-DDRS_REF_3D=%d" % (1 if is3d else 0),
           "-DDRS_REF_STREAMING=%d" % streaming, "-DDRS_REF_MX=%d" % mx, "-DDRS_REF_MY=%d" % my,
           "-DDRS_REF_STEP=%d" % step, "-o", so, os.path.join(HERE, "ref_wrap.cu")]
    rc, out = sh(cmd)
    if rc != 0:
        raise RuntimeError("nvcc failed on %s:\n%s" % (case, out))
    return dict(stencil=stem, is3d=is3d, L=L, M=M, N=N, iterations=iters, step=step, options=opts)


def build(only=None, verbose=True):
    if not os.path.isdir(REF):
        print("build_ref: %s absent -- keeping prebuilt oracle/_ref" % REF)
        return False
    generator()
    meta = {}
    meta_path = os.path.join(OUT, "cases.json")
    if os.path.exists(meta_path):
        meta = json.load(open(meta_path))
    for case, (stem, is3d, (L, M, N), iters, step, opts) in CASES.items():
        if only and case not in only:
            continue
        so = os.path.join(OUT, "libref_%s.so" % case)
        if os.path.exists(so) and case in meta and meta[case].get("options") == opts:
            continue
        meta[case] = build_one(case, stem, is3d, (L, M, N), iters, step, opts, os.path.join(OUT, "cases", case), so)
        meta[case]["so"] = "libref_%s.so" % case
        meta[case]["cu"] = "cases/%s/%s.cu" % (case, stem)
        if verbose:
            print("build_ref: built", case)
        json.dump(meta, open(meta_path, "w"), indent=1, sort_keys=True)
    return True


if __name__ == "__main__":
    build(only=set(sys.argv[1:]) or None)
