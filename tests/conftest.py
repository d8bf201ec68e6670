import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from cpu_checkers import Oracle, build_checkers
    build_checkers()
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from cpu_checkers import Ref, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref/libhrd_ref.so not built (no /root/reference here)")
    return Ref()
