import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_py import Oracle, build_oracle
    build_oracle()
    return Oracle()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_rect_xsbl.npz"))
    return {k: g[k] for k in g.files}


@pytest.fixture(scope="session")
def cv_golden():
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "cv2_bm_golden.npz"))
    return {k: g[k] for k in g.files}


@pytest.fixture(scope="session")
def libpath():
    """Builds libu96stereo.so in-tree if needed (nvcc cross-compiles without a GPU)."""
    from u96_slam_b200 import build
    return build.build()
