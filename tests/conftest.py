import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the product never builds itself (uf3_b200._native.lib() fails loudly without the library);
    # the TEST session does, so that a fresh checkout can run `pytest` directly
    lib = os.path.join(ROOT, "uf3_b200", "lib", "libuf3b.so")
    if not os.path.isfile(lib):
        import subprocess
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "uf3_b200", "csrc"), "-j", str(os.cpu_count() or 4)],
                       check=True)


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
