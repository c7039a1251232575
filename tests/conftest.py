import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _seeded():
    """Every test draws its random inputs from the same seeded global generators: a tolerance either holds or it
    does not -- no run-to-run flakiness on the GPU box."""
    torch.manual_seed(1234)
    np.random.seed(1234)
    yield


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    return {k: (torch.from_numpy(z[k]) if z[k].ndim > 0 else z[k].item()) for k in z.files}


@pytest.fixture
def golden():
    return load_golden
