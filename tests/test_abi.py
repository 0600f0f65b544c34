"""The C-ABI library builds, loads and exports every symbol include/vrag_b200.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_cuda


def test_library_exports_every_declared_symbol():
    from verbatim_rag_b200 import _native
    lib = _native.load_library()
    header = open(os.path.join(ROOT, "include", "vrag_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(vrag_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    assert sorted(_native.EXPORTS) == declared
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"
    assert b"sm_100a" in lib.vrag_version()


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    from verbatim_rag_b200 import _native
    with pytest.raises(_native.NativeError) as ei:
        _native.Context(0)
    assert ei.value.code == _native.VRAG_ERR_CUDA
    assert "no CPU fallback" in str(ei.value)
    from verbatim_rag_b200 import B200VectorStore
    with pytest.raises(_native.NativeError):
        B200VectorStore(dense_dim=768)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "verbatim_rag_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports oracle"


def test_stale_library_is_not_loaded_silently(monkeypatch):
    """A .so built from other sources than the tree's (build.sha256 mismatch) must be rebuilt or refused, never
    dlopen'ed with this module's argtypes."""
    import pytest
    from verbatim_rag_b200 import _native, build
    _native.load_library()
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(build, "_digest", lambda: "0" * 64)
    with pytest.raises(_native.NativeError, match="stale"):
        _native.load_library(build_if_missing=False)
