"""GPU tests of the frame pipeline (jxlb_pipeline_*): many frames in flight through N decoder contexts with a bounded
number of heavy-stage slabs. Every frame that comes out is compared with the oracle bit for bit, in every output mode,
including the BASELINE sizes (3840x2160 and 7680x4320 synthetic frames: the layouts the bench decodes)."""
import ctypes

import numpy as np
import pytest

import bench
from conftest import fixture_bytes

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method="thread")]


@pytest.fixture(scope="module")
def pipe():
    import jxl_oxide_b200 as J
    p = J.Pipeline(0, workers=6, heavy_frames=2)
    yield p
    p.close()


def _oracle_planar(oracle, data, threads=16):
    img = oracle.OracleImage(data, threads=threads)
    want = img.frame(0)[0]
    img.close()
    return want


def _as_array(addr, nbytes, dtype, shape):
    buf = (ctypes.c_uint8 * nbytes).from_address(addr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()


def test_mixed_frames_planar_f32(pipe, oracle):
    """Frames of different sizes and codings interleaved, more frames than workers, 2 heavy slots: every output is the
    oracle's, whatever order the frames finish in."""
    names = ["opsin_inverse", "bike", "issue_311", "grayalpha", "cafe", "upsampling", "squeeze_edge", "minecraft_vardct_e7"]
    datas = [fixture_bytes(n, "input.jxl") for n in names] + [bench.synth_frame(2000, 1500, 3), bench.synth_frame(1000, 600, 7)]
    want = [_oracle_planar(oracle, d) for d in datas]
    for rep in range(2):
        for i, d in enumerate(datas):
            pipe.submit(data=d, mode=pipe.OUT_PLANAR_F32, tag=100 * rep + i)
    seen = set()
    while pipe.in_flight:
        tag, addr, nbytes = pipe.wait(want_output=True)
        w = want[tag % 100]
        assert nbytes == w.nbytes
        got = _as_array(addr, nbytes, np.float32, w.shape)
        pipe.release_output(addr)
        assert np.array_equal(got.view(np.uint32), w.view(np.uint32)), f"frame {tag} differs from the oracle"
        seen.add(tag)
    assert len(seen) == 2 * len(datas)


def test_caller_buffers_and_u8(pipe, oracle):
    import jxl_oxide_b200 as J
    data = bench.synth_frame(1000, 600, 7)
    d = J.Decoder(0)
    d.decode(data)
    want_u8 = d.frame_to_buffer(0, np.uint8)
    want_u16 = d.frame_to_buffer(0, np.uint16)
    want_f32 = d.frame_planar(0)
    d.close()
    out_u8 = np.zeros_like(want_u8)
    out_f32 = np.zeros_like(want_f32)
    pipe.submit(data=data, out=out_u8, tag=1)
    pipe.submit(data=data, out=out_f32, tag=2)
    pipe.submit(data=data, mode=pipe.OUT_U16, tag=3)
    pipe.preload(5, data)
    pipe.submit(slot=5, tag=4)  # decode only
    got16 = None
    while pipe.in_flight:
        tag, addr, nbytes = pipe.wait(want_output=True)
        if tag == 3:
            got16 = _as_array(addr, nbytes, np.uint16, want_u16.shape)
            pipe.release_output(addr)
        elif tag == 4:
            assert addr is None
    assert np.array_equal(out_u8, want_u8)
    assert np.array_equal(out_f32.view(np.uint32), want_f32.view(np.uint32))
    assert np.array_equal(got16, want_u16)


def test_errors_are_per_frame(pipe, oracle):
    """A corrupt frame reports its own error; frames around it are unaffected and the slab / buffer rings stay usable."""
    import jxl_oxide_b200 as J
    good = bench.synth_frame(1000, 600, 7)
    want = _oracle_planar(oracle, good)
    bad = bytearray(good)
    for i in range(len(bad) // 2, len(bad) // 2 + 64):
        bad[i] ^= 0x5a
    bad = bytes(bad)
    trunc = good[: len(good) // 3]
    for rep in range(3):
        pipe.submit(data=good, mode=pipe.OUT_PLANAR_F32, tag=10 + rep)
        pipe.submit(data=bad, mode=pipe.OUT_PLANAR_F32, tag=20 + rep)
        pipe.submit(data=trunc, mode=pipe.OUT_PLANAR_F32, tag=30 + rep)
    ok = failed = 0
    while pipe.in_flight:
        try:
            tag, addr, nbytes = pipe.wait(want_output=True)
        except J.JxlError:
            failed += 1
            continue
        got = _as_array(addr, nbytes, np.float32, want.shape)
        pipe.release_output(addr)
        if 10 <= tag < 20:
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
            ok += 1
    assert ok == 3 and failed >= 3  # the truncated stream always fails; the corrupted one fails or decodes to other pixels
    # still healthy
    pipe.submit(data=good, mode=pipe.OUT_PLANAR_F32, tag=99)
    tag, addr, nbytes = pipe.wait(want_output=True)
    got = _as_array(addr, nbytes, np.float32, want.shape)
    pipe.release_output(addr)
    assert tag == 99 and np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("size,seed,extra", [((3840, 2160), 1, ()), ((7680, 4320), 1, ()),
                                             ((7680, 4320), 2, ("--distance", "2.0", "--epf-iters", "3")),
                                             ((3840, 2160), 1, ("--modular",))],
                         ids=["synth4k_d1", "synth8k_d1", "synth8k_d2_epf3", "synthmod4k"])
def test_baseline_sizes_match_oracle(pipe, oracle, size, seed, extra):
    """The bench workloads themselves: final planes and the integer HF coefficients against the oracle."""
    import jxl_oxide_b200 as J
    w, h = size
    dist = 1.0
    ex = list(extra)
    if "--distance" in ex:
        i = ex.index("--distance")
        dist = float(ex[i + 1])
        del ex[i:i + 2]
    data = bench.synth_frame(w, h, seed, distance=dist, extra=tuple(ex))
    modular = "--modular" in ex
    img = oracle.OracleImage(data, threads=32, capture=not modular)
    want = img.frame(0)[0]
    want_coeff = [] if modular else img.stage("hf_coeff", np.int32)
    img.close()
    pipe.submit(data=data, mode=pipe.OUT_PLANAR_F32, tag=7)
    pipe.submit(data=data, mode=pipe.OUT_PLANAR_F32, tag=8)
    while pipe.in_flight:
        tag, addr, nbytes = pipe.wait(want_output=True)
        got = _as_array(addr, nbytes, np.float32, want.shape)
        pipe.release_output(addr)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "pipeline output differs from the oracle"
    if modular:
        return
    d = J.Decoder(0)
    d.set_capture(True)
    d.decode(data)
    for g, wc in zip(d.stage("hf_coeff", np.int32), want_coeff):
        assert np.array_equal(g, wc), "HF coefficients differ from the oracle"
    d.close()
