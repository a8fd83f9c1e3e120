"""CPU-side checks of the drop-in boundary: libkmsc.so loads and exports every
symbol include/kmsc.h declares; without a GPU compute entry points fail loudly
(there is no CPU fallback). No kernels are launched here."""
import ctypes as C
import re
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200"))


@pytest.fixture(scope="module")
def so():
    import kmsc
    if not kmsc.LIB_PATH.exists():
        kmsc.build()
    return C.CDLL(str(kmsc.LIB_PATH))


def test_header_symbols_exported(so):
    header = (ROOT / "include" / "kmsc.h").read_text()
    names = set(re.findall(r"\b(kmsc_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 25
    for n in sorted(names):
        assert hasattr(so, n), f"{n} declared in include/kmsc.h but not exported"


def test_binding_lists_every_symbol():
    import kmsc
    header = (ROOT / "include" / "kmsc.h").read_text()
    names = set(re.findall(r"\b(kmsc_[a-z0-9_]+)\s*\(", header))
    assert names == set(kmsc.EXPORTS)


def test_fails_loudly_without_gpu(so):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    so.kmsc_ctx_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    rc = so.kmsc_ctx_create(0, None, C.byref(h))
    so.kmsc_last_error.restype = C.c_char_p
    assert rc != 0 and b"no CPU fallback" in so.kmsc_last_error()


def test_product_does_not_touch_oracle():
    """the shipped package must not import, link or call anything under oracle/"""
    pkg = ROOT / "kmer-sets-compression_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")) \
            + list(pkg.rglob("*.cc")) + list(pkg.rglob("Makefile")):
        text = p.read_text()
        assert "kmsc_oracle" not in text and "oracle/" not in text and "_oracle" not in text, p
