import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (test infrastructure): C restatement of the reference, see oracle/oracle.h"""
    from oracle import binding
    binding.build()
    return binding.Oracle(threads=max(1, min(8, os.cpu_count() or 1)))


@pytest.fixture(scope="session")
def ml():
    """the product: python mirror over the C ABI of libmultilinear_b200.so"""
    import multilinear_b200
    multilinear_b200.build()
    return multilinear_b200.api
