import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def oracle():
    import pb_oracle
    pb_oracle.build()
    return pb_oracle


@pytest.fixture(scope='session')
def ctx():
    from peppan_b200._lib import Context
    c = Context(0)
    yield c
    c.close()
