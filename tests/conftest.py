import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def gpu_lib():
    """Builds (if needed) and loads libbandup_gpu.so; no fallback of any kind."""
    from bandup_b200 import build
    build.build()
    from bandup_b200 import _capi
    return _capi.load()


@pytest.fixture()
def unfolder(gpu_lib):
    from bandup_b200.unfold import GpuUnfolder
    u = GpuUnfolder(0)
    yield u
    u.close()
