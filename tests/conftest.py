import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # Cosmology objects build their tables with the device builders (the product path).  A test session on a
    # machine without a GPU prepares oracle inputs with the numpy twins instead -- explicitly, here: the product
    # default stays "device" and fails loudly without a CUDA device.
    import ccl_b200
    from ccl_b200 import _lib
    try:
        ngpu = _lib.load().ccl_b200_device_count()
    except Exception:
        ngpu = 0
    if ngpu < 1:
        ccl_b200.set_table_builder("host")


@pytest.fixture(scope="session")
def synth():
    """One synthetic LSST-Y1-like table set shared by the tests."""
    from ccl_b200 import synthetic as S
    bg = S.background()
    pk = S.power(bg)
    subs, colls = S.lsst_like_tracers(bg, with_ia=True, with_rsd=True, with_mag=True)
    return dict(bg=bg, pk=pk, subs=subs, colls=colls)
