"""The C-ABI library loads on a CPU-only box and exports every symbol include/amr_b200.h declares
(no compute calls here; the product has no CPU path)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    txt = (ROOT / "include" / "amr_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(amr_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    from blastamr_b200 import _build, capi
    lib = ctypes.CDLL(str(_build.build_cuda()))
    names = declared_symbols()
    assert len(names) >= 20
    for nm in names:
        assert hasattr(lib, nm), f"{nm} declared in amr_b200.h but not exported"
    assert set(names) == set(capi.EXPORTS), "capi.py binding and header disagree"


def test_no_gpu_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from blastamr_b200 import capi
    with pytest.raises(capi.AmrError):
        capi.Context(0)


def test_product_does_not_import_oracle():
    """only tests/, smoke() and bench.py's cpu legs may touch oracle/ (building it in _build.py is not using it)"""
    for py in (ROOT / "blastamr_b200").rglob("*.py"):
        txt = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), py
        if py.name != "_build.py":
            assert "liboracle" not in txt and "amr_oracle" not in txt, py
    for src in (ROOT / "blastamr_b200" / "csrc").glob("*"):
        txt = src.read_text()
        assert "amr_oracle" not in txt and "liboracle" not in txt and "#include \"../../oracle" not in txt, src
