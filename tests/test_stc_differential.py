"""Randomised differential test of the .stc front end against the reference generator
(oracle/_ref/drstencil_ref, built from /root/reference/main.cpp): random descriptions (keys in any
order, unknown tokens such as the shipped "iteratioins" typo, duplicate keys and points, radius 1-3,
2D and 3D, assorted coefficient spellings) run through both command lines with random --step/--dist;
exit code, first output line and the L/M/N/Iterations/Halo macros of the emitted programs must agree
(/root/reference/drstencil_2d.hpp:48-97, drstencil.hpp:52-103)."""
import os
import random
import re
import subprocess

import pytest

from helpers import ROOT

CLI = os.path.join(ROOT, "drstencil_b200", "bin", "drstencil")
REF = os.path.join(ROOT, "oracle", "_ref", "drstencil_ref")


def random_stc(rng, is3d):
    r = rng.choice([1, 1, 2, 3])
    dim = 3 if is3d else 2
    # the largest positive slow-axis offset bounds every other offset (what Halo is derived from)
    pts = [tuple([r] + [0] * (dim - 1))]
    pts += [tuple(rng.randint(-r, r) for _ in range(dim)) for _ in range(rng.randint(2, 9))]
    if rng.random() < 0.7:
        pts.append(tuple([-r] + [0] * (dim - 1)))
    coef = lambda: rng.choice(["0.2", "0.1", "0.05", ".5", "1e-1", "2", "-0.1", "0.125", "0.3333333", "1.5e-2", "0.07", "3"])
    keys = []
    if is3d or rng.random() < 0.2:
        keys.append(("L", str(rng.choice([64, 96, 130]))))
    keys.append(("M", str(rng.choice([64, 128, 512, 1000]))))
    keys.append(("N", str(rng.choice([64, 128, 512, 1024]))))
    has_iterations = rng.random() < 0.9
    if has_iterations:
        keys.append(("iterations", str(rng.choice([2, 4, 6, 10]))))
    rng.shuffle(keys)
    lines = []
    for name, value in keys:
        lines.append("%s %s" % (name, value))
        if rng.random() < 0.15:
            lines.append("%s %s" % (rng.choice(["foo", "iteratioins", "#", "bar"]), rng.choice(["7", "x"])))
        if rng.random() < 0.1:
            lines.append("%s %s" % (name, value))
    text = "\n".join(lines) + "\nstencil\n"
    for p in pts:
        text += " ".join(map(str, p)) + " " + coef() + rng.choice(["\n", " ", "\t\n"])
    return text, has_iterations


def _defines(path):
    if not os.path.exists(path):
        return {}
    return {m.group(1): int(m.group(2)) for m in re.finditer(r"^#define\s+(\w+)\s+(-?\d+)", open(path).read(), re.M)}


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/drstencil_ref not built")
def test_random_descriptions_parse_and_analyse_like_the_reference(built, tmp_path):
    compared = emitted = 0
    for seed in range(160):
        rng = random.Random(seed)
        is3d = rng.random() < 0.4
        text, has_iterations = random_stc(rng, is3d)
        (tmp_path / "t.stc").write_text(text)
        step, dist = rng.choice([1, 1, 2, 3]), rng.choice([0, 0, 1, 2])
        argv = (["--3d"] if is3d else []) + ["--step", str(step)] + (["--dist", str(dist)] if dist else []) + \
            ["--bx", "64", "--by", "16"]
        for f in ("o.cu", "r.cu"):
            if os.path.exists(tmp_path / f):
                os.remove(tmp_path / f)
        ours = subprocess.run([CLI] + argv + ["-o", "o.cu", "t.stc"], cwd=tmp_path, stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True, timeout=60)
        try:
            ref = subprocess.run([REF] + argv + ["-o", "r.cu", "t.stc"], cwd=tmp_path, stdout=subprocess.PIPE,
                                 stderr=subprocess.STDOUT, text=True, timeout=5)
        except subprocess.TimeoutExpired:
            continue
        compared += 1
        assert ours.returncode == ref.returncode, (seed, argv, text, ours.stdout, ref.stdout)
        assert ours.stdout.strip().splitlines()[:1] == ref.stdout.strip().splitlines()[:1], (seed, argv, text)
        if ref.returncode != 0:
            continue
        emitted += 1
        mine, theirs = _defines(tmp_path / "o.cu"), _defines(tmp_path / "r.cu")
        pairs = {"M": "GridM", "N": "GridN", "Halo": "Halo"}
        if is3d:
            pairs["L"] = "GridL"
        if has_iterations:                 # without the key the reference prints an uninitialised int
            pairs["Iterations"] = "Iterations"
        for r, o in pairs.items():
            assert theirs.get(r) == mine.get(o), (seed, argv, r, theirs.get(r), mine.get(o), text)
    assert compared >= 150 and emitted >= 60, (compared, emitted)
