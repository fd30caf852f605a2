"""The C-ABI shared library loads without a GPU and exports every symbol include/ccl_b200.h
declares; compute entries fail loudly (CCL_B200_ERROR_CUDA) when no device is present.
No compute call is made on a GPU here -- that is the job of the `-m gpu` parity tests."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from ccl_b200 import _lib
    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ccl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ccl_b200_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from ccl_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "libccl_b200.so does not export %s" % n
    # the ctypes table binds exactly the declared set
    assert sorted(_lib.SIGNATURES) == names


def test_header_cites_reference_interfaces():
    text = open(os.path.join(ROOT, "include", "ccl_b200.h")).read()
    for cite in ("src/ccl_cls.c:265-363", "include/ccl_cls.h:20-26", "pyccl/ccl_cls.i:98-112",
                 "src/ccl_tracers.c:569-691", "src/ccl_f2d.c:92-201", "pyccl/ccl_pk2d.i:17-28"):
        assert cite in text


def test_default_params_match_reference_defaults(lib):
    # src/ccl_core.c:67-141
    from ccl_b200 import _lib
    p = _lib.Params()
    lib.ccl_b200_default_params(C.byref(p))
    assert (p.K_MAX, p.K_MIN, p.DLOGK_INTEGRATION) == (1e3, 5e-5, 0.025)
    assert (p.INTEGRATION_LIMBER_EPSREL, p.N_ITERATION, p.A_SPLINE_MIN) == (1e-4, 1000, 0.1)
    # the Gauss-Kronrod entries hold GSL's rule key like the reference: GSL_INTEG_GAUSS41 = 4 (src/ccl_core.c:40)
    assert p.INTEGRATION_LIMBER_GAUSS_KRONROD_POINTS == 4
    import ccl_b200 as ccl
    assert ccl.gsl_params.INTEGRATION_LIMBER_GAUSS_KRONROD_POINTS == 4
    assert b"sm_100a" in lib.ccl_b200_version()


def test_no_gpu_fails_loudly(lib):
    """Without a device the constructors set CCL_B200_ERROR_CUDA: there is no CPU fallback."""
    if lib.ccl_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from ccl_b200 import _lib
    st = C.c_int(0)
    h = lib.ccl_b200_cosmology_new(None, C.byref(st))
    assert not h and st.value == _lib.ERROR_CUDA
    assert "no CPU fallback" in _lib.last_error()
    st = C.c_int(0)
    b = lib.ccl_b200_batch_new(2, None, C.byref(st))
    assert not b and st.value == _lib.ERROR_CUDA
    import ccl_b200 as cb
    with pytest.raises(cb.CCLB200Error):
        cb.LibBatch(1)


def test_status_is_inout_noop(lib):
    """pyccl/ccl.i:34 convention: a call is a no-op when *status != 0 on entry."""
    st = C.c_int(1234)
    h = lib.ccl_b200_batch_new(1, None, C.byref(st))
    assert not h and st.value == 1234


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under ccl_b200/ may reference it."""
    pkg = os.path.join(ROOT, "ccl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "liblimber_oracle" not in src, f
