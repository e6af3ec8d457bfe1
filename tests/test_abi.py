"""CPU-only checks of the C-ABI boundary: the library builds, loads, exports every symbol the
header declares, and fails loudly (no CPU fallback) when there is no CUDA device."""

import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    with open(os.path.join(REPO, "include", "haphic_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hh_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    import __graft_entry__ as g
    g.build()
    from haphic_b200 import _lib
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    assert sorted(_lib.exported_symbols()) == syms, "ctypes signature table out of sync with the header"
    assert lib.hh_version() == 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from haphic_b200._lib import Context, HHError
    with pytest.raises(HHError) as e:
        Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_package_does_not_import_oracle():
    """The product path must never route through the oracle."""
    pkg = os.path.join(REPO, "haphic_b200")
    for root, _d, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                with open(os.path.join(root, fn)) as f:
                    src = f.read()
                assert "import oracle" not in src and "from oracle" not in src, fn
