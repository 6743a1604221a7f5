import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def reference_fixtures():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "reference_fixtures.npz")))


@pytest.fixture(scope="session")
def oracle_golden():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "oracle_golden.npz")))


def tolerance(y):
    """Stated fp32 tolerance of the CUDA path against the float64 oracle:
    |yhat_gpu - yhat_ref| <= 1e-4 * max|y| + 1e-3 per element."""
    import numpy as np
    return 1e-4 * float(np.nanmax(np.abs(np.where(np.isfinite(y), y, 0.0)))) + 1e-3
