import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)')


@pytest.fixture(scope='session')
def port_oracle():
    from oracle.cpu_oracle import Oracle, build
    build()
    return Oracle('port')


@pytest.fixture(scope='session')
def ref_oracle():
    """The unmodified reference kernels run on the CPU (oracle/_ref, built where /root/reference exists)."""
    from oracle.cpu_oracle import Oracle, available, build
    build()
    if not available('reference'):
        pytest.skip('oracle/_ref/libgendr_ref_cpu.so not built (needs /root/reference)')
    return Oracle('reference')
