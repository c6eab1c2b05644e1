import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Build what is missing (the driver normally runs __graft_entry__.build() first)."""
    need = [os.path.join(ROOT, "prototype2_b200", "libprtc.so"), os.path.join(ROOT, "prototype2_b200", "libprtscene.so"),
            os.path.join(ROOT, "oracle", "_build", "libprt_l1.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()
    yield
