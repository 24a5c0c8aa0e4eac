"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/ghnd_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from hnd_ghnd_object_detectors_b200 import build
    return build.build_library()


def _declared():
    text = open(os.path.join(ROOT, "include", "ghnd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ghnd_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in ghnd_b200.h is not exported" % n


def test_python_binding_covers_header(lib_path):
    from hnd_ghnd_object_detectors_b200 import _lib
    assert sorted(_lib.SIGNATURES.keys()) == _declared()
    lib = _lib.load()
    assert lib.ghnd_abi_version() == 1
    assert isinstance(_lib.last_error(), str)


def test_argument_validation_without_gpu(lib_path):
    """Invalid arguments are rejected on the host before any CUDA call."""
    from hnd_ghnd_object_detectors_b200 import _lib
    lib = _lib.load()
    assert lib.ghnd_quantize_u8(None, 16, 8, 0, None, None, None, 0, None) == 1
    assert "null" in _lib.last_error()
    d = _lib.ConvDesc()
    h = ctypes.c_void_p()
    d.kind, d.N, d.H, d.W, d.C, d.K, d.R, d.S, d.stride, d.pad = 0, 1, 8, 8, 3, 64, 2, 2, 1, 1
    assert lib.ghnd_conv_plan_create(ctypes.byref(d), ctypes.byref(h)) == 1
    assert "multiples of 64" in _lib.last_error()


def test_no_cpu_fallback_in_product():
    """The product package must not import or call the oracle (tests-only infrastructure)."""
    pkg = os.path.join(ROOT, "hnd_ghnd_object_detectors_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), fn
