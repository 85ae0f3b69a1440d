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


def test_new_entry_points_reject_bad_arguments_without_a_device():
    """Argument validation needs no GPU: a NULL decoder is an error value, never a crash."""
    import jxl_oxide_b200
    L = ctypes.CDLL(_built_lib())
    L.jxlb_set_hf_streams_per_cta.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    L.jxlb_frame_write_to_device.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_size_t]
    assert L.jxlb_set_hf_streams_per_cta(None, 64) == jxl_oxide_b200.ERR_INVALID_ARG
    buf = ctypes.create_string_buffer(16)
    assert L.jxlb_frame_write_to_device(None, 0, 0, 0, buf, 16) == jxl_oxide_b200.ERR_INVALID_ARG


def test_round2_entry_points_reject_bad_arguments_without_a_device():
    """Pipeline, allocation budget, section decode and upsample entry points: error values for bad arguments / no GPU."""
    import torch
    import jxl_oxide_b200 as J
    L = J.load_library()
    h = ctypes.c_void_p()
    assert L.jxlb_decoder_create_ex(0, 1 << 20, None) == J.ERR_INVALID_ARG
    assert L.jxlb_pipeline_create(0, None, None) == J.ERR_INVALID_ARG
    assert L.jxlb_decode_frame_sections(None, b"x", 1, None, 0, None) == J.ERR_INVALID_ARG
    assert L.jxlb_upsample(None, None, 1, 1, 1, 2, None, 2) == J.ERR_INVALID_ARG
    assert L.jxlb_decode_hf_groups(None, b"x", 1, None, 0, None, None) == J.ERR_INVALID_ARG
    assert L.jxlb_dequant_idct(None, b"x", 1, None, 0, None, None) == J.ERR_INVALID_ARG
    assert L.jxlb_modular_decode_groups(None, b"x", 1, None, 0, 0, None, None, 0) == J.ERR_INVALID_ARG
    assert L.jxlb_pipeline_submit(None, None, 0, 0, 0, None, 0, 0) == J.ERR_INVALID_ARG
    assert L.jxlb_pipeline_release_output(None, None) == J.ERR_INVALID_ARG
    assert L.jxlb_pipeline_workers(None) == 0 and L.jxlb_pipeline_launch_count(None) == 0
    if not torch.cuda.is_available():
        assert L.jxlb_decoder_create_ex(0, 0, ctypes.byref(h)) == J.ERR_CUDA and not h.value
        assert L.jxlb_pipeline_create(0, None, ctypes.byref(h)) == J.ERR_CUDA and not h.value
