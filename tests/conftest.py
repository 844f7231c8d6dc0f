import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Build oracle + CUDA library once (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def room02():
    from mav_active_3d_planning_b200 import synth
    return synth.room_map(0.2)


@pytest.fixture(scope="session")
def room01():
    from mav_active_3d_planning_b200 import synth
    return synth.room_map(0.1)


CAM_BASE = dict(resolution_x=320, resolution_y=240, focal_length=160.0, ray_length=5.0)
CAM_CFG4 = dict(resolution_x=640, resolution_y=480, focal_length=320.0, ray_length=8.0)
