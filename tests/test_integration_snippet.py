"""INTEGRATION.md's patch to the reference host, as a compilable C unit
(examples/reference_host_patch.c): it must compile as C against
include/gravity_cuda.h and link against libgravity_cuda; on a GPU it must then
reproduce the reference's two_body_circular bits."""
import os
import subprocess

import pytest

import proj_entropy_b200 as g
from conftest import ROOT, load_fixture

SRC = os.path.join(ROOT, "examples", "reference_host_patch.c")
PKG = os.path.join(ROOT, "proj_entropy_b200")


def build(tmp_path):
    exe = str(tmp_path / "host_patch")
    cmd = ["gcc", "-std=gnu11", "-O2", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), "-o", exe, SRC,
           "-L", PKG, "-lgravity_cuda", "-Wl,-rpath," + PKG]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    assert p.stderr.strip() == "", p.stderr          # warning-free C
    return exe


def test_patch_compiles_and_links_as_c(tmp_path):
    g.load_cuda_library()
    exe = build(tmp_path)
    if g.device_count() < 1:
        # no GPU here: the shim must refuse to run rather than fall back
        p = subprocess.run([exe, "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert p.returncode == 1 and "no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_patched_host_reproduces_reference_bits(tmp_path):
    exe = build(tmp_path)
    fx = load_fixture("ref_two_body_circular")
    cp = [c for c in fx["checkpoints"] if c["steps"] == 100][0]
    p = subprocess.run([exe, "100"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    rows = [l.split() for l in p.stdout.splitlines()]
    for i, r in enumerate(rows):
        got = [float.fromhex(v) for v in r[1:]]
        assert got == [cp["x_f"][i], cp["y_f"][i], cp["vx_f"][i], cp["vy_f"][i]]
