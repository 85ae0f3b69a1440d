"""GPU tests of the round-2 boundary additions: allocation budget (AllocTracker), jxlb_upsample on device buffers,
jxlb_decode_frame_sections (a frame handed over as separate TOC-section buffers)."""
import numpy as np
import pytest

import bench
from conftest import fixture_bytes

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]


def test_upsample_entry_matches_oracle(oracle):
    """features::upsample (4x) of the `upsampling` conformance stream's filtered planes, on caller-owned device memory."""
    import torch
    import jxl_oxide_b200 as J
    img = oracle.OracleImage(fixture_bytes("upsampling", "input.jxl"), threads=4, capture=True)
    src = img.stage("epf", np.float32)
    want = img.stage("upsampled", np.float32)
    img.close()
    d = J.Decoder(0)
    for s, w in zip(src, want):
        got = d.upsample(torch.from_numpy(np.ascontiguousarray(s)).cuda(), 4).cpu().numpy()
        assert got.shape == w.shape
        assert np.array_equal(got.view(np.uint32), w.view(np.uint32))
    d.close()


def test_allocation_budget(oracle):
    import jxl_oxide_b200 as J
    data = bench.synth_frame(2000, 1500, 3)
    want = oracle.OracleImage(data, threads=8).frame(0)[0]
    small = J.Decoder(0, mem_limit=8 << 20)          # 2000x1500 needs ~70 MB of planes
    with pytest.raises(J.JxlError) as e:
        small.decode(data)
    assert e.value.code == J.ERR_OUT_OF_MEMORY
    tiny = bench.synth_frame(520, 392, 11)           # fits the budget; the failed decode left nothing behind
    small.decode(tiny)
    assert np.array_equal(small.frame_planar(0).view(np.uint32), oracle.OracleImage(tiny, threads=4).frame(0)[0].view(np.uint32))
    small.close()
    big = J.Decoder(0, mem_limit=1 << 30)
    big.decode(data)
    assert np.array_equal(big.frame_planar(0).view(np.uint32), want.view(np.uint32))
    big.close()


def test_decode_frame_sections(oracle):
    import jxl_oxide_b200 as J
    data = bench.synth_frame(1000, 600, 7)
    want = oracle.OracleImage(data, threads=8).frame(0)[0]
    cuts = [0, 300, 301, 5000, 5000, len(data) // 2, len(data)]
    header, sections = data[:cuts[1]], [data[a:b] for a, b in zip(cuts[1:-1], cuts[2:])]
    d = J.Decoder(0)
    d.decode_sections(header, sections)
    assert np.array_equal(d.frame_planar(0).view(np.uint32), want.view(np.uint32))
    d.close()


def test_stage_entry_points_vardct(oracle):
    """jxlb_decode_hf_groups / jxlb_dequant_idct: the seams between entropy decode, dequant + IDCT and the filters, on
    caller-owned device planes, against the same stages of the oracle."""
    import jxl_oxide_b200 as J
    data = bench.synth_frame(1000, 600, 5)
    img = oracle.OracleImage(data, threads=8, capture=True)
    want_coeff, want_idct, want_px = img.stage("hf_coeff", np.int32), img.stage("idct", np.float32), img.frame(0)[0]
    img.close()
    d = J.Decoder(0)
    coeff = d.decode_hf_groups(data)
    assert len(coeff) == 3
    for g, w in zip(coeff, want_coeff):
        assert tuple(g.shape) == w.shape and np.array_equal(g.cpu().numpy(), w)
    for g, w in zip(d.dequant_idct(data), want_idct):
        assert tuple(g.shape) == w.shape and np.array_equal(g.cpu().numpy().view(np.uint32), w.view(np.uint32))
    assert d.modular_decode_groups(data) == []       # a VarDCT frame without extra channels codes no Modular channel
    d.decode(data)                                   # the decoder is left usable
    assert np.array_equal(d.frame_planar(0).view(np.uint32), want_px.view(np.uint32))
    d.close()


def test_stage_entry_point_modular(oracle):
    import jxl_oxide_b200 as J
    data = bench.synth_frame(600, 400, 3, extra=("--modular",))
    img = oracle.OracleImage(data, threads=8, capture=True)
    want = img.stage("modular_coded", np.int32)
    img.close()
    d = J.Decoder(0)
    got = d.modular_decode_groups(data)
    assert len(got) == len(want) and len(got) > 3    # squeezed: residual channels besides the three colour channels
    for g, w in zip(got, want):
        assert tuple(g.shape) == w.shape and np.array_equal(g.cpu().numpy(), w)
    with pytest.raises(J.JxlError) as e:
        d.decode_hf_groups(data)                     # no HF groups in a Modular frame
    assert e.value.code == J.ERR_UNSUPPORTED
    d.close()
