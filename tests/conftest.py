import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    # a fresh on-disk beam-table cache per test session: the tables under test are the ones this build produces
    import tempfile

    os.environ["OK_BEAM_CACHE_DIR"] = tempfile.mkdtemp(prefix="ok_beam_cache_")
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU check")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the product library and the oracle once per session (no-ops when up to date)."""
    import __graft_entry__ as g

    g.build()
    yield


def have_gpu() -> bool:
    try:
        import openkitchen_b200 as ok

        e = ok.Env(device=0)
        e.close()
        return True
    except Exception:
        return False
