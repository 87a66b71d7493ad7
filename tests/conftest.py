import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import loader
    return loader.load_oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle import loader
    ref = loader.load_reference()
    if ref is None:
        pytest.skip("oracle/_ref/libavif_ref.so not built (needs /root/reference)")
    return ref


@pytest.fixture(scope="session")
def product():
    import libavif_b200
    return libavif_b200.lib()


@pytest.fixture(scope="session")
def gpu_product(product):
    import libavif_b200
    if not libavif_b200.device_available():
        pytest.fail("GPU test selected but libavif_b200 sees no CUDA device")
    return product
