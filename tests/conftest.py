import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gpu_renderer_factory():
    import ifrit_v2_b200  # noqa: F401
    from ifrit_v2_b200 import renderer as R
    made = []

    def make():
        r = R.TileRasterRendererCuda(0)
        r.Init()
        made.append(r)
        return r
    yield make
    for r in made:
        r.close()
