"""The C-ABI library loads and exports every symbol include/gqec_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge

    ge.build()
    from graphqec_b200 import backend

    return backend.load_library()


def _declared():
    text = open(os.path.join(ROOT, "include", "gqec_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gq_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = _declared()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gqec_b200.h but not exported"


def test_binding_covers_header(lib):
    from graphqec_b200 import backend

    assert sorted(backend.SIGNATURES) == _declared()


def test_version_and_error_channel(lib):
    assert lib.gq_version() >= 100
    assert isinstance(lib.gq_last_error(), bytes)


def test_argument_validation_without_gpu(lib):
    # NULL out-pointer is rejected before any CUDA call
    rc = lib.gq_dem_create(0, 0, 0, None, None, None, None, None, None, None, None, None, 0, None)
    assert rc == -1
    assert b"out is NULL" in lib.gq_last_error()


def test_no_device_is_a_loud_error(lib):
    if lib.gq_device_count() > 0:
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = lib.gq_dem_create(1, 1, 0, None, None, None, None, None, None, None, None, None, 0, ctypes.byref(h))
    assert rc == -2 and b"no CPU fallback" in lib.gq_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "graphqec_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
