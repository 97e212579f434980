import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # quiet the reference-style info logging
    try:
        from loguru import logger
        logger.remove()
        logger.add(sys.stderr, level="WARNING")
    except Exception:
        pass


@pytest.fixture(scope="session")
def goldens():
    with open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")) as fh:
        return json.load(fh)["entries"]


@pytest.fixture(scope="session")
def built_library():
    """The in-tree shared library (built on demand; nvcc cross-compiles without a GPU)."""
    from chiron_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def cuda_device(built_library):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
