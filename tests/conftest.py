import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "photons-2.0_gpu-p2p-redundancy_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

DEMO_BOX = 100000.0
DEMO_NSIDE = 32
DEMO_MASS = 211.75382579190332
THETA = 0.4


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """The CPU suite needs the oracle .so and the product .so files (nvcc cross-compiles here)."""
    import __graft_entry__ as ge
    ge.build()


@pytest.fixture(scope="session")
def demo_pos():
    """Positions of the reference's bundled demo IC (1_Indexing/demo/ic_lcdm.gdt2): 32^3 particles,
    float32 on disk, widened to float64 (fixture made by tests/golden/make_golden.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "demo_lcdm_pos_f32.npy")).astype(np.float64)


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "demo_lists.json")) as f:
        return json.load(f)
