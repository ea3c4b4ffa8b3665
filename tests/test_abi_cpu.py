"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/openems_b200.h declares, and the product path fails loudly (no CPU fallback) when no
GPU is present.  No compute call is made here."""
import ctypes
import os
import re

import pytest

import openems_b200
from openems_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    txt = open(os.path.join(ROOT, header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(oems_(?:cuda|synth)_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = openems_b200.load_library()
    names = declared_symbols("include/openems_b200.h") + declared_symbols("openems_b200/csrc/host/synthetic_operator.h")
    assert len(names) > 50
    raw = ctypes.CDLL(openems_b200.library_path())
    for n in names:
        assert hasattr(raw, n), "missing export: " + n
    # and every declared symbol has a ctypes signature in the Python binding
    missing = [n for n in names if n not in _lib.SIGNATURES]
    assert not missing, missing
    assert L.oems_cuda_abi_version() == 1


def test_coeff_entry_layout_matches_header():
    assert ctypes.sizeof(_lib.CoeffEntry) == 128
    assert _lib.CoeffEntry.pml.offset == 48 and _lib.CoeffEntry.pml_ii.offset == 88


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "openems_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inc")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    op = openems_b200.Operator_CUDA((8, 8, 8))
    with pytest.raises(openems_b200.EngineError) as ei:
        op.CreateEngine()
    assert "no usable CUDA device" in str(ei.value) or "CUDA" in str(ei.value)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(openems_b200.LibraryNotBuilt):
        _lib.load_library()
