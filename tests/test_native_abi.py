"""CPU: the C-ABI library builds, loads, and exports every symbol include/b200lopq.h declares; the
product has no CPU fallback (creating a handle without a GPU fails loudly)."""
import os
import re

import pytest

from columbiaimagesearch_b200 import _native, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _native.load_library()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200lopq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2l_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libb200lopq.so does not export %s" % n
    # the ctypes table covers the whole header, nothing more, nothing less
    assert sorted(_native.SIGNATURES) == names


def test_version(lib):
    assert lib.b2l_version() == _native.ABI_VERSION


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_native.NativeError) as e:
        _native.Handle(0)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_oracle():
    """Nothing under the package may import oracle/ (the checker)."""
    pkg = os.path.join(ROOT, "columbiaimagesearch_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith(".py"):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
