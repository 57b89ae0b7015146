import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests"), os.path.join(REPO, "sahs-deformable-nerf_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """After a GPU session, print the field kernels' diagnostic word (readable even if a kernel trapped)."""
    mod = sys.modules.get("sahs_b200.lib")
    if mod is None or getattr(mod, "_lib", None) is None:
        return
    import ctypes as C
    out = (C.c_int * 4)()
    mod._lib.sahs_field_status(out)
    if out[0] != 0:
        print(f"\n[sahs] field kernel status word: code={out[0]} tag={out[1]} block={out[2]} thread={out[3]}")
