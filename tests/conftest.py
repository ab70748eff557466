import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib():
    """librecur_b200.so through ctypes (the product)."""
    from recur_b200 import api
    if not os.path.exists(api.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return api.load_library()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference compiled in place (IEEE build)."""
    import oracle
    if not oracle.have_ref():
        if os.path.exists("/root/reference/recur-nn.c"):
            oracle.build(ref=True, port=False)
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return oracle.load_ref(strict=True)


@pytest.fixture(scope="session")
def ref_fast():
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    return oracle.load_ref(strict=False)


@pytest.fixture(scope="session")
def port():
    import oracle
    return oracle.load_port()


@pytest.fixture(scope="session")
def gpu_lib(lib):
    if lib.rnn_b200_device_count() < 1:
        pytest.skip("no CUDA device")
    return lib
