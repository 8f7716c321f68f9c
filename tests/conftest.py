import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


os.environ.setdefault("VSD_WATCHDOG_S", "120")   # libvideosd: abort with a message if the device stops answering (never hang the run)
# operator tests: the library fills a GEMM's split-K workspace and dense output with NaN bytes before every launch, so a tile
# a configuration fails to write cannot hide behind what an earlier configuration of the same shape left in the buffer
os.environ.setdefault("VSD_POISON", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run on the B200 box with -m gpu")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))


@pytest.fixture(scope="session")
def oracle_models():
    """Seeded random-init oracle UNet / TAESD (fp32, CPU) shared by the session."""
    from oracle.weights import build_taesd, build_unet

    return build_unet(), build_taesd()
