"""CPU-only checks of the drop-in boundary: the library loads, exports exactly what
include/bgls_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import os
import re

import pytest

from bgls_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "bgls_b200.h")).read()
    return sorted(set(re.findall(r"\b(bgls_[a-z0-9_]+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert _declared() == sorted(_native.EXPORTS)


def test_library_exports_every_declared_symbol():
    L = _native.load()
    for name in _declared():
        assert hasattr(L, name), name
    assert b"sm_100a" in L.bgls_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_native.BglsError):
        _native.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "bgls_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
