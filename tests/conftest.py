import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session', autouse=True)
def _build_oracle():
    from oracle import oracle
    oracle._lib()  # compiles the C restatement on first use (seconds)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


ROUTE_CASES = ['small', 'substeps', 'shuffled', 'chain']


@pytest.fixture(params=ROUTE_CASES)
def route_golden(request):
    return load_golden(f'route_{request.param}.npz')


def require_cuda():
    import river_route_b200 as rr
    if not rr.cuda_available():
        pytest.fail('gpu-marked test started without a usable CUDA device (there is no CPU fallback)')
