import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def models():
    from nanocall_b200 import models as M
    return {m["name"]: m for m in M.load_builtin_models()}


@pytest.fixture(scope="session")
def port():
    import oracle_lib
    return oracle_lib.port()


@pytest.fixture(scope="session")
def ref():
    import oracle_lib
    if not oracle_lib.have_ref():
        pytest.skip("oracle/_ref/libncref.so not built and /root/reference absent")
    return oracle_lib.ref()


@pytest.fixture(scope="session")
def ctx():
    from nanocall_b200 import api
    c = api.Context(0)
    yield c
    c.close()
