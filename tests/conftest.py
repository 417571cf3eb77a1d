import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Compile the CUDA library and the C oracle once per session (no-op when up to date)."""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def orc():
    import aus_oracle
    return aus_oracle


@pytest.fixture(scope="session")
def aus():
    import aus_rust_b200
    return aus_rust_b200


@pytest.fixture(scope="session")
def ctx(aus):
    return aus.Context(0)
