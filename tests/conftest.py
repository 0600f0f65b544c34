"""pytest config: the ``gpu`` marker and import paths.

``-m "not gpu"`` runs everywhere (oracle vs golden vectors, host logic, C-ABI exports);
``-m gpu`` needs a B200 and calls the CUDA path through the C-ABI.
/root/reference exists only in the build container: tests that import the reference's
own Python (plugin conformance) skip when it is absent and never run under ``-m gpu``.
"""
import os
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "reference: imports the reference's Python from /root/reference")


def _install_rapidfuzz_stub():
    """verbatim_core.extractors imports rapidfuzz at module import (extractors.py:18); it is not
    installed offline.  Only the LLM extractor's fuzzy verifier uses it -- out of scope here."""
    if "rapidfuzz" in sys.modules:
        return
    try:
        import rapidfuzz  # noqa: F401
        return
    except Exception:
        pass
    rf = types.ModuleType("rapidfuzz")
    fz = types.ModuleType("rapidfuzz.fuzz")
    fz.partial_ratio = lambda a, b, **k: 0.0
    fz.ratio = lambda a, b, **k: 0.0
    fz.partial_ratio_alignment = lambda a, b, **k: None
    rf.fuzz = fz
    sys.modules["rapidfuzz"] = rf
    sys.modules["rapidfuzz.fuzz"] = fz


@pytest.fixture(scope="session")
def reference_pkgs():
    """Import the reference's own packages (build container only)."""
    if not os.path.isdir(REFERENCE):
        pytest.skip("/root/reference not present on this machine")
    _install_rapidfuzz_stub()
    for p in (os.path.join(REFERENCE, "packages", "core"), REFERENCE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import verbatim_core  # noqa: F401
    import verbatim_rag  # noqa: F401
    return sys.modules["verbatim_rag"], sys.modules["verbatim_core"]


def has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
