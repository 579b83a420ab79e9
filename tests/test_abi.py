"""The C-ABI library loads on a CPU-only box and exports exactly what include/puresound_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from puresound_b200 import _lib, build

    build.build()
    return _lib.load()


def _declared():
    text = open(os.path.join(ROOT, "include", "puresound_b200.h")).read()
    return sorted(set(re.findall(r"PS_API[^;(]*?\b(ps_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from puresound_b200 import _lib

    names = _declared()
    assert len(names) >= 29
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_struct_mirrors(lib):
    from puresound_b200 import _lib

    assert lib.ps_version() == 1
    for which, st in _lib.STRUCTS.items():
        assert lib.ps_struct_size(which) == ctypes.sizeof(st)
    assert lib.ps_struct_size(99) == -1
    assert lib.ps_error_string(0) == b"ok"
    assert b"invalid" in lib.ps_error_string(-1)


def test_slot_counts(lib):
    assert lib.ps_gemm_stats_slots(3999, 512) == 32 * 4
    assert lib.ps_gemm_stats_slots(1, 1) == 1
    assert lib.ps_dwconv_stats_slots(3999, 512) == 256  # max(streaming 250, tiled 16 x 16)
    assert lib.ps_dwconv_stats_slots(50, 24) == 4


def test_argument_validation_without_gpu(lib):
    """NULL descriptors / bad sizes are rejected before any CUDA call is made."""
    from puresound_b200 import _lib

    assert lib.ps_gemm(None, None) == -1
    d = _lib.GemmDesc()
    assert lib.ps_gemm(ctypes.byref(d), None) == -1
    assert lib.ps_dwconv(None, None) == -1
    assert lib.ps_lstm(None, None) == -1
    assert lib.ps_ola(None, 1, 1, 1, 1, None, 0, None, None) == -1
    with pytest.raises(ValueError):
        _lib.check(-1, "x")
    with pytest.raises(NotImplementedError):
        _lib.check(-2, "x")


def test_integration_doc_struct_mirror_matches_library(lib):
    """The ps_gemm_t ctypes mirror printed in INTEGRATION.md (what a reference maintainer would paste) has the size and
    the field order of the real descriptor - a short struct would make the kernel read past it."""
    from puresound_b200 import _lib

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    body = text[text.index("class ps_gemm_t(C.Structure)"):]
    body = body[body.index("_fields_ = ["):body.index("]\n") + 1]
    ns = {"C": ctypes}
    exec(body.strip(), ns)  # noqa: S102  (a literal list of (name, ctype) pairs from our own document)
    doc_fields = ns["_fields_"]

    class Mirror(ctypes.Structure):
        _fields_ = doc_fields

    assert ctypes.sizeof(Mirror) == lib.ps_struct_size(0)
    assert [n for n, _ in doc_fields] == [n for n, _ in _lib.GemmDesc._fields_]
    for (n, t), (_, t2) in zip(doc_fields, _lib.GemmDesc._fields_):
        assert ctypes.sizeof(t) == ctypes.sizeof(t2), n
