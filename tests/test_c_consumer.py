"""include/drstencil.h is plain C: a C99 consumer (examples/host_loop.c, the reference's emitted host
loop written against the ABI) compiles with gcc, links libdrstencil.so and runs -- without a GPU up
to the loud DRS_E_NOGPU, on the GPU through the whole loop with the gold check."""
import os
import subprocess

import pytest

from helpers import ROOT


def _build(tmp_path, name="host_loop"):
    exe = str(tmp_path / name)
    lib = os.path.join(ROOT, "drstencil_b200")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", name + ".c"), "-L", lib, "-ldrstencil", "-Wl,-rpath," + lib, "-o", exe],
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


def test_c99_slab_consumer_builds_and_fails_loudly_without_gpu(built, tmp_path):
    """examples/slab_loop.c: the multi-GPU run loop (drs_run_slab) is reachable from plain C."""
    import drstencil_b200 as drs
    exe = _build(tmp_path, "slab_loop")
    if drs.lib().drs_device_count() == 0:
        r = subprocess.run([exe, os.path.join(ROOT, "stc", "3d7pt_star.stc"), str(tmp_path)], stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, timeout=120)
        assert r.returncode == 1 and "drs_set_device(local) -> -7" in r.stdout, r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("step", ["1", "2"])
def test_c99_slab_consumer_under_torchrun(built, tmp_path, step):
    """One C process per GPU (torchrun --no-python), handles traded through files: slab run == whole-grid run,
    bit for bit, one kernel launch per sweep."""
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    exe = _build(tmp_path, "slab_loop")
    rdv = tmp_path / ("rdv" + step)
    rdv.mkdir()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--no-python", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", "29613", exe,
                        os.path.join(ROOT, "stc", "3d7pt_star.stc"), str(rdv), step],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.count("SLAB_LOOP_OK") == world, r.stdout[-3000:]
