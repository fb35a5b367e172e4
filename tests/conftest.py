import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Outputs of the unmodified reference on oracle/recipes.py inputs (tests/golden/make_golden.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def state_dicts():
    from oracle import recipes
    cache = {}

    def get(dataset):
        if dataset not in cache:
            cache[dataset] = recipes.make_state_dict(recipes.DATASETS[dataset]["n_embed"], seed=0)
        return cache[dataset]
    return get
