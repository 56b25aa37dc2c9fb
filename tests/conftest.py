import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # GPU tests fail loudly (not skip) on a box without a GPU only when explicitly selected with -m gpu
    pass


@pytest.fixture(scope="session")
def coco_names():
    return [f"class{i}" for i in range(80)]
