#!/usr/bin/env python
"""Generates tests/golden/ref_emitted.json by running the REAL reference generator
(oracle/_ref/drstencil_ref, built by oracle/build_ref.py from /root/reference/main.cpp) in the
authoring container.  The reference ships no golden vectors of its own, so
these are outputs of the reference itself, captured here and committed:

  for every shipped stencil x step x dist option: exit code, the macros it emits
  (L M N Iterations Range Halo Dist) and the gold expression -- term order, offsets and the
  coefficient literal text.

Usage (authoring container only; the GPU box has no /root/reference):
    python oracle/build_ref.py && python tests/golden/make_golden.py
"""
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
GEN = os.path.join(ROOT, "oracle", "_ref", "drstencil_ref")
REF = "/root/reference"

STENCILS = ["2d5pt_star", "2d5pt_cross", "2d9pt_star", "2d9pt_box", "2d9pt_cross", "2d25pt_box",
            "3d7pt_star", "3d9pt_cross"]
# (step, dist option, extra) combinations to record per stencil
COMBOS = [(1, 0, []), (1, 2, []), (2, 0, []), (2, 1, []), (3, 0, []), (1, 0, ["--merge-forward", "1"]),
          (2, 0, ["--merge-forward", "1"]), (2, 3, ["--merge-forward", "9"])]
DEEP = {"2d5pt_star": [4], "2d9pt_box": [4], "3d7pt_star": [3]}

TERM = re.compile(r"\(([^)]+)\) \* in((?:\[[kji][+-]?\d*\])+)")
IDX = re.compile(r"\[([kji])([+-]?\d*)\]")


def parse_emitted(text, is3d):
    macros = {}
    for m in re.finditer(r"^#define (L|M|N|Iterations|Range|Halo|Dist|Bx|By|Sn) (-?\d+)\s*$", text, re.M):
        macros[m.group(1)] = int(m.group(2))
    g0 = text.index("__global__ void gold_")
    body = text[g0:text.index("int main", g0)]
    expr = body[body.index("out["):]
    expr = expr[:expr.index(";")]
    terms = []
    for m in TERM.finditer(expr):
        off = {"k": 0, "j": 0, "i": 0}
        for ax, v in IDX.findall(m.group(2)):
            off[ax] = int(v) if v not in ("", "+", "-") else 0
        terms.append([off["k"], off["j"], off["i"], m.group(1)])
    return macros, terms


def main():
    if not os.path.exists(GEN):
        sys.exit("build oracle/_ref first: python oracle/build_ref.py")
    out = {}
    for stem in STENCILS:
        is3d = stem.startswith("3d")
        combos = list(COMBOS) + [(s, 0, []) for s in DEEP.get(stem, [])]
        for step, dist, extra in combos:
            if is3d and step > 3:
                continue
            with tempfile.TemporaryDirectory() as td:
                stc = os.path.join(td, stem + ".stc")
                with open(os.path.join(REF, "benchmarks", stem, stem + ".stc")) as f:
                    open(stc, "w").write(f.read())
                args = [GEN] + (["--3d"] if is3d else []) + ["--step", str(step)]
                if dist:
                    args += ["--dist", str(dist)]
                args += extra + ["--streaming", "--bx", "256", "--sn", "64", "--check", "-o", "o.cu", stem + ".stc"]
                r = subprocess.run(args, cwd=td, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                key = "%s|step=%d|dist=%d|%s" % (stem, step, dist, " ".join(extra))
                rec = {"rc": r.returncode, "stdout": r.stdout.strip()}
                cu = os.path.join(td, "o.cu")
                if r.returncode == 0 and os.path.exists(cu):
                    macros, terms = parse_emitted(open(cu).read(), is3d)
                    rec["macros"] = macros
                    rec["gold_terms"] = terms
                out[key] = rec
    path = os.path.join(ROOT, "tests", "golden", "ref_emitted.json")
    json.dump(out, open(path, "w"), indent=0, sort_keys=True)
    print("wrote %s: %d records" % (path, len(out)))


if __name__ == "__main__":
    main()
