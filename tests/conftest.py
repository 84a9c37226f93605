import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "emul")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def cfg2():
    """BASELINE config 2 inputs (reverse parking, N=80, 3 obstacles), 8 problems, numpy default_rng(0)."""
    from obca_b200.scenarios import reverse_parking_batch
    return reverse_parking_batch(8, 80, 0)
