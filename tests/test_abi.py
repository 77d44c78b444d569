"""The drop-in boundary: both shared libraries load and export every function
that include/*.h declares; without a GPU the compute library refuses to work
instead of falling back."""
import ctypes
import os
import re

import pytest

import proj_entropy_b200 as g
from conftest import ROOT


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gravity_[a-z0-9_]+)\s*\(", text)))


def test_cuda_library_exports_every_declared_symbol():
    lib = g.load_cuda_library()
    names = declared_functions("gravity_cuda.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(lib._declared) == names          # the ctypes mirror is complete too


def test_host_library_exports_every_declared_symbol():
    lib = g.load_host_library()
    names = declared_functions("gravity_host.h")
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(lib._declared) == names


def test_abi_is_plain_c():
    # no C++ mangled gravity symbols leak out of the boundary
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "proj_entropy_b200", "libgravity_cuda.so")],
                         stdout=subprocess.PIPE, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    declared = set(declared_functions("gravity_cuda.h"))
    assert declared <= set(exported)
    # an entry point that lost its extern "C" would show up mangled (_Z...gravity_cuda_...)
    assert [s for s in exported if s.startswith("_Z") and "gravity_cuda_" in s] == []
    # and nothing unmangled is exported under the library's prefix without being in the header
    assert {s for s in exported if s.startswith("gravity_")} == declared


def test_no_cpu_fallback_without_gpu():
    if g.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(g.GravityError) as ei:
        g.GravityCuda(16)
    assert "no CPU fallback" in str(ei.value)
    with pytest.raises(g.GravityError):
        g.fp64_peak(0)


def test_product_does_not_reference_the_oracle():
    # the oracle is test infrastructure: nothing under the package or include/
    # may import, link or name it
    for base in ("proj_entropy_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert "oracle" not in text.replace("no oracle", ""), os.path.join(dirpath, f)
    ldd = os.popen("ldd %s" % os.path.join(ROOT, "proj_entropy_b200", "libgravity_cuda.so")).read()
    assert "oracle" not in ldd


def test_torch_can_be_imported_after_the_cuda_library():
    """Both link libnccl.so.2 and a process holds one library per soname: with the system's NCCL in
    first, torch's CUDA library would miss symbols of the newer NCCL it ships with.  The ctypes
    mirror therefore loads that bundled NCCL first; either import order must work."""
    import subprocess
    import sys
    for code in ("import proj_entropy_b200 as g; g.load_cuda_library(); import torch; print(torch.__version__)",
                 "import torch; import proj_entropy_b200 as g; g.load_cuda_library(); print(torch.__version__)"):
        p = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                           timeout=300, cwd=ROOT)
        assert p.returncode == 0, p.stderr[-1500:]
