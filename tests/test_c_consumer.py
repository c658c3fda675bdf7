"""include/drstencil.h is plain C: a C99 consumer (examples/host_loop.c, the reference's emitted host
loop written against the ABI) compiles with gcc, links libdrstencil.so and runs -- without a GPU up
to the loud DRS_E_NOGPU, on the GPU through the whole loop with the gold check."""
import os
import subprocess

import pytest

from helpers import ROOT


def _build(tmp_path):
    exe = str(tmp_path / "host_loop")
    lib = os.path.join(ROOT, "drstencil_b200")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "host_loop.c"), "-L", lib, "-ldrstencil", "-Wl,-rpath," + lib, "-o", exe],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return exe


def test_c99_consumer_builds_and_fails_loudly_without_gpu(built, tmp_path):
    import drstencil_b200 as drs
    exe = _build(tmp_path)
    r = subprocess.run([exe, os.path.join(ROOT, "stc", "2d9pt_box.stc"), "4", "512", "512"], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=120)
    assert "Halo 1 Dist 1" in r.stdout and "kernel dr_2d9pt_box, 4 timestep(s) per sweep" in r.stdout, r.stdout
    if drs.lib().drs_device_count() == 0:
        assert r.returncode == 0 and "no GPU: drs_run_host -> -7" in r.stdout, r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("stc,step", [("2d9pt_box", "4"), ("2d5pt_star", "1"), ("3d7pt_star", "2")])
def test_c99_consumer_runs_the_reference_loop_on_gpu(built, tmp_path, stc, step):
    exe = _build(tmp_path)
    size = ["96", "200"] if stc.startswith("3d") else ["520", "648"]
    r = subprocess.run([exe, os.path.join(ROOT, "stc", stc + ".stc"), step] + size, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert "GPU finished computing." in r.stdout and "[Test] RMS Error:" in r.stdout
