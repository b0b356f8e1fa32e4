import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # checker libraries: the port always builds; the reference only where /root/reference exists
    # (the GPU box uses the prebuilt oracle/_ref/*.so that travelled with the snapshot)
    try:
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except Exception:
        pass


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def ref():
    import dsvlibs
    if not dsvlibs.have_ref():
        pytest.skip("oracle/_ref/libdsv1ref.so not built (needs /root/reference)")
    return dsvlibs.ref()


@pytest.fixture(scope="session")
def port():
    import dsvlibs
    return dsvlibs.port()


@pytest.fixture(scope="session")
def gpu():
    import dsvlibs
    return dsvlibs.gpu()
