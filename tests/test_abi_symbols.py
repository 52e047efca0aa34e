"""The C-ABI library loads on a CPU-only box and exports every symbol include/skani_b200.h declares; the product
refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "skani_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from skani_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    L = _lib.load()
    names = declared_symbols()
    assert len(names) >= 25
    bound = {n for n, _, _ in _lib.SYMBOLS}
    for n in names:
        assert hasattr(L, n), "not exported: " + n
        assert n in bound, "no ctypes signature for " + n


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import skani_b200 as sk
    with pytest.raises(sk.host.SkaniError):
        sk.Context(0)


def test_result_struct_layout_matches_header():
    from skani_b200._lib import AniResult, MapParams, SketchParams
    assert C.sizeof(AniResult) == 72 and C.sizeof(SketchParams) == 12 and C.sizeof(MapParams) == 40
