import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_sessionstart(session):
    """Built artefacts are kept out of git: from a clean tree, build the CUDA library (nvcc
    cross-compiles without a GPU) and the checker before collecting, as __graft_entry__.build()
    does."""
    import shutil
    import subprocess

    lib = os.path.join(ROOT, "cunumeric_b200", "libcunumeric_b200.so")
    if not os.path.exists(lib) and shutil.which("nvcc") and shutil.which("make"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "cunumeric_b200", "csrc"),
                        "-j", str(os.cpu_count() or 8)], check=False)
    ref = os.path.join(ROOT, "oracle", "_ref", "libcunumeric_ref.so")
    if not os.path.exists(ref) and os.path.isdir("/root/reference/src/cunumeric"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j", "4"], check=False)


def _gpu_available() -> bool:
    try:
        from cunumeric_b200 import _lib

        return _lib.load().cnb_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not silently pass: only skip GPU tests when
    # they were not explicitly selected.
    if "gpu" in (config.getoption("-m") or ""):
        return
    if _gpu_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
