import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, GOLDEN):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_sessionstart(session):
    """The native pieces are git-ignored build artefacts: build them when a fresh checkout is tested directly
    (what `__graft_entry__.build()` does; nvcc cross-compiles without a GPU)."""
    import glob
    lib = os.path.join(ROOT, "grakel_b200", "libgrakel_b200.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    have_packer = bool(glob.glob(os.path.join(ROOT, "grakel_b200", "_fastpack*.so")))
    if not (os.path.exists(lib) and have_packer) and os.path.exists(nvcc):
        subprocess.run(["bash", os.path.join(ROOT, "grakel_b200", "csrc", "build.sh")], check=False)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
