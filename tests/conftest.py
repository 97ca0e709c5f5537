import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _ensure_built():
    """The .so files are built artefacts (git-ignored); build them on first use so that a fresh checkout can run the suite."""
    from hmc_b200 import build as b
    if not os.path.exists(b.SO):
        b.build()
    if not os.path.exists(b.HOST_SO) or not os.path.exists(b.CLI):
        b.build_host()


def pytest_configure(config):
    _ensure_built()
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "ref: needs the reference sources at /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    skip_ref = pytest.mark.skip(reason="/root/reference not present")
    has_ref = os.path.isdir("/root/reference")
    for it in items:
        if "gpu" in it.keywords and not has_gpu:
            it.add_marker(skip_gpu)
        if "ref" in it.keywords and not has_ref:
            it.add_marker(skip_ref)
