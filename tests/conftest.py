import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mf-nerf_b200")):
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


@pytest.fixture(scope="session")
def ref_vren():
    """The reference's own vren kernels (oracle/_ref/vren_ref*.so, built by oracle/build_ref_vren.sh), or None."""
    import glob
    import importlib.util
    import torch  # noqa: F401  (must be imported before the extension)
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "vren_ref*.so"))
    if not so or not torch.cuda.is_available():
        return None
    spec = importlib.util.spec_from_file_location("vren_ref", so[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
