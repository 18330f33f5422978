import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(GOLDEN / name, allow_pickle=False))


def unpack_masks(fx):
    shape = tuple(int(v) for v in fx["masks_shape"])
    return np.unpackbits(fx["masks"])[: int(np.prod(shape))].reshape(shape)


@pytest.fixture(scope="session")
def body9():
    return np.array([[-0.22, 0, 0, -0.26, 0, 0, -0.1704612, 0.4309841, -0.00670862]])
