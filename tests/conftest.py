import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def _cuda_device_present():
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a machine without a GPU skips the gpu-marked tests instead of failing them."""
    if not any("gpu" in it.keywords for it in items) or _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device (gpu-marked tests run on the B200 box)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle harness (builds liboracle.so on first use)."""
    from oracle import harness
    harness.build()
    return harness


@pytest.fixture(scope="session")
def pkg():
    """The product package `meshing.jl_b200` (loaded by path: the directory name has a dot)."""
    import subprocess
    from __graft_entry__ import PKG_DIR, load_package
    lib = os.path.join(PKG_DIR, "lib", "libb200iso.so")
    if not os.path.exists(lib):  # fresh checkout: build the CUDA library in-tree (nvcc cross-compiles without a GPU)
        subprocess.check_call(["make", "-C", os.path.join(PKG_DIR, "csrc")], stdout=subprocess.DEVNULL)
    return load_package()


@pytest.fixture
def c_example(pkg, tmp_path):
    """examples/extract.c compiled as C99 against include/b200iso.h and linked with the built library."""
    import subprocess
    exe = str(tmp_path / "extract")
    libdir = os.path.join(ROOT, "meshing.jl_b200", "lib")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O2", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "extract.c"), "-o", exe, "-L" + libdir, "-lb200iso", "-lm", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe
