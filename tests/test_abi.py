"""CPU test: the C-ABI library loads (no GPU needed, no compute calls) and exports every symbol include/sfgwas_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sfgwas_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sfg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from sfgwas_b200 import _lib
    from sfgwas_b200.build import build

    build()  # no-op when up to date (nvcc cross-compiles without a GPU)
    L = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/sfgwas_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == syms
    assert _lib.load().sfg_version() >= 1


def test_no_cpu_fallback_without_device(gpu_available):
    """Without a CUDA device context creation must fail loudly (there is no CPU path)."""
    if gpu_available:
        pytest.skip("a GPU is present")
    from sfgwas_b200 import CryptoParams, SfgError

    with pytest.raises(SfgError):
        CryptoParams(8, [0x1FFFEC001], [0x800004001], 2.0 ** 30)


def test_product_never_imports_the_oracle():
    """The product package must not reference oracle/ (only tests, smoke and bench's baseline legs may)."""
    pkg = os.path.join(ROOT, "sfgwas_b200")
    for dp, _, fs in os.walk(pkg):
        if os.path.basename(dp) in ("build", "lib", "__pycache__"):
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "sfg_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, os.path.join(dp, f)
