import os
import sys
import warnings

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    """GPU tests must fail loudly -- never silently skip -- when selected with -m gpu on a
    box where the CUDA library or device is missing; under -m 'not gpu' they are deselected."""
    return


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def cuda_lib():
    """The loaded C-ABI library; building it first if this checkout has no .so yet."""
    from pytenet_b200 import _build, _lib
    if _build.needs_build():
        _build.build()
    return _lib.load()


@pytest.fixture(autouse=True)
def _quiet_lanczos_breakdown():
    # the breakdown warning fires at every chain edge -- expected
    with warnings.catch_warnings():
        warnings.filterwarnings("ignore", message="beta\\[", category=RuntimeWarning)
        yield
