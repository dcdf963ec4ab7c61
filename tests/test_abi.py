"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/qq_b200.h declares, and the product
path fails loudly (no CPU fallback) when no B200 is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "qq_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qq_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(pkg):
    lib = ctypes.CDLL(pkg.lib_path())
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "libqq_b200.so does not export %s" % s


def test_binding_covers_header(pkg):
    from quisquis_rust_b200.binding import EXPORTS
    assert set(EXPORTS) == set(declared_symbols())


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; this test checks the no-GPU failure mode")
    with pytest.raises(pkg.QQError, match="no CPU fallback"):
        pkg.Engine(0)


def test_product_does_not_reference_oracle():
    """Nothing under quisquis-rust_b200/ may import, link or call the oracle."""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "quisquis-rust_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".inc", ".hpp", ".h", ".cpp")):
                t = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"(import\s+ristretto_ref|import\s+c_oracle|libqq_oracle|qq_oracle\.c|#include\s+\"[^\"]*oracle)", t):
                    bad.append(f)
    assert not bad, bad
