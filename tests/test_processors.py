"""Processor routines behind the C ABI: the host-only ones run here without a GPU against the
oracle and the closed forms derivable from the reference source; the GPU resize / full image preprocessing are `-m gpu`."""
import numpy as np
import pytest

from aha_b200 import processors as P, synth
from oracle import audio as OA
from oracle import qwen3vl as OV


def test_img_smart_resize_matches_the_oracle_and_known_answers():
    assert P.img_smart_resize(1080, 1920) == (1088, 1920)            # round(1080 / 32) = 34
    assert P.img_smart_resize(2048, 2048) == (2048, 2048)
    rng = np.random.default_rng(0)
    for _ in range(300):
        h, w = int(rng.integers(8, 6000)), int(rng.integers(8, 6000))
        if max(h, w) // min(h, w) > 200:
            continue
        mn, mx = int(rng.choice([3136, 65536, 262144])), int(rng.choice([1048576, 16777216]))
        assert P.img_smart_resize(h, w, 32, mn, mx) == OV.img_smart_resize(h, w, 32, mn, mx), (h, w, mn, mx)
    with pytest.raises(P.ProcessorError, match="aspect ratio"):
        P.img_smart_resize(10, 5000)


def test_placeholder_expansion():
    ids = np.array([5, 9, 7, 9, 3, 9], np.uint32)
    assert P.expand_placeholders(ids, 9, [3, 1]).tolist() == [5, 9, 9, 9, 7, 9, 3, 9]
    assert P.expand_placeholders(ids, 9, [0]).tolist() == [5, 7, 9, 3, 9]
    assert np.array_equal(P.expand_placeholders(ids, 9, [2, 2, 2]), OV.expand_placeholders(ids, 9, [2, 2, 2]))
    # a 1080p image: 2040 copies of <|image_pad|>; 30 s of audio: 390 copies of <|audio_pad|>
    assert (P.expand_placeholders([1, 2, 3], 2, [68 * 120 // 4]) == 2).sum() == 2040
    assert (P.expand_placeholders([1, 2, 3], 2, [P.feat_extract_output_length(3000)]) == 2).sum() == 390


def test_audio_helpers():
    for n in list(range(0, 1300)) + [3000, 2999, 120000]:
        assert P.feat_extract_output_length(n) == OA.get_feat_extract_output_lengths(n) == synth.asr_audio_tokens(n)
    assert [P.feat_extract_output_length(n) for n in (1, 100, 250, 3000)] == [1, 13, 33, 390]
    rng = np.random.default_rng(1)
    for scale in (0.3, 1.0, 2.5):
        x = (rng.standard_normal(4000) * scale).astype(np.float32)
        assert np.array_equal(P.float_range_normalize(x), OA.float_range_normalize(x[None])[0])
    assert np.array_equal(P.float_range_normalize(np.zeros(8, np.float32)), np.zeros(8, np.float32))
    for total, sec in ((16000 * 30, 1200.0), (16000 * 2500, 1200.0), (16000 * 2400, 1200.0), (12345, 0.5)):
        assert P.split_audio_into_chunks(total, 16000, sec) == OA.split_audio_into_chunks(total, 16000, sec)
    assert P.split_audio_into_chunks(16000 * 2400, 16000, 1200.0) == [19200000, 19200000, 0]      # the reference pushes the empty remainder too


def test_sinc_resample_bank_and_oracle_properties():
    """The filter bank built by the library (host f32) against the oracle's restatement, and closed-form properties of the resampler:
    a constant stays constant away from the edges, a low-frequency sine keeps its frequency and amplitude, lengths follow ceil(new * n / orig)."""
    for orig, new in ((48000, 16000), (44100, 16000), (8000, 16000), (22050, 16000), (24000, 16000), (16000, 24000)):
        taps, width = P.sinc_resample_bank(orig, new)
        g = int(np.gcd(orig, new))
        want, w2 = OA.get_sinc_resample_kernel(orig, new, g)
        assert width == w2 and taps.shape == want.shape == (new // g, 2 * width + orig // g)
        assert np.abs(taps - want).max() <= 2e-7          # libm cos / sin against numpy's, a few ulp of values <= 1
    x = np.ones((1, 4800), np.float32)
    y = OA.resample_simple(x, 48000, 16000)
    assert y.shape == (1, 1600) and np.abs(y[0, 40:-40] - 1.0).max() < 2e-3
    t = np.arange(44100, dtype=np.float64) / 44100.0
    y = OA.resample_simple(np.sin(2 * np.pi * 440.0 * t).astype(np.float32)[None], 44100, 16000)
    assert y.shape == (1, 16000)
    ref = np.sin(2 * np.pi * 440.0 * np.arange(16000) / 16000.0)
    assert np.abs(y[0, 100:-100] - ref[100:-100]).max() < 5e-3
    assert OA.resample_simple(x, 16000, 16000).shape == x.shape
    assert OA.resample_simple(np.zeros((1, 1001), np.float32), 44100, 16000).shape == (1, int(np.ceil(160 * 1001 / 441)))
    with pytest.raises(P.ProcessorError, match="Frequencies must be positive"):
        P.sinc_resample_bank(0, 16000)


def test_oracle_resize_properties():
    """The CatmullRom restatement: identity at equal size, constant images stay constant, a 2x box-like downscale of a ramp stays monotone."""
    img = synth.synth_image(40, 56, 3)
    assert np.array_equal(OV.resize_exact_catmullrom(img, 40, 56), img)
    flat = np.full((33, 47, 3), 137, np.uint8)
    assert np.all(OV.resize_exact_catmullrom(flat, 64, 96) == 137) and np.all(OV.resize_exact_catmullrom(flat, 16, 20) == 137)
    ramp = np.repeat(np.arange(0, 240, 2, dtype=np.uint8)[None, :, None], 24, 0).repeat(3, 2)
    small = OV.resize_exact_catmullrom(ramp, 12, 60)
    assert small.shape == (12, 60, 3) and np.all(np.diff(small[0, :, 0].astype(int)) >= 0)


@pytest.mark.gpu
def test_gpu_resize_and_image_preprocess_match_the_oracle():
    from conftest import make_model
    cfg, w, m = make_model("qwen3vl", "tiny", max_ctx=1024, max_patches=4096)
    try:
        for (h, wd, nh, nw) in ((100, 150, 128, 160), (300, 200, 96, 64), (77, 91, 77, 91), (64, 64, 352, 352), (480, 640, 96, 160)):
            img = synth.synth_image(h, wd, h + wd)
            got = m.image_resize(img, nh, nw)
            want = OV.resize_exact_catmullrom(img, nh, nw)
            assert np.array_equal(got, want), (h, wd, nh, nw, int(np.abs(got.astype(int) - want.astype(int)).max()))
        # whole Qwen3VLProcessor image path on a size that needs the resize: 250 x 333 -> img_smart_resize -> (256, 320)... and a tiny one that is upscaled
        for (h, wd) in ((250, 333), (90, 70)):
            img = synth.synth_image(h, wd, 11)
            pv, grid = m.image_preprocess(img)
            want_pv, want_grid = OV.process_image(img)
            assert grid.tolist() == want_grid.tolist() and pv.shape == want_pv.shape
            assert float(np.abs(pv - want_pv).max()) <= 1e-6
    finally:
        m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("orig,new,n", [(48000, 16000, 48000), (44100, 16000, 30011), (8000, 16000, 4001), (22050, 16000, 22050), (16000, 16000, 100),
                                        (24000, 16000, 7), (16000, 24000, 1), (44100, 16000, 0)])
def test_gpu_resample_matches_the_oracle(orig, new, n):
    """aha_b200_resample (filter bank on the host, strided convolution on the GPU) against the restated resample_simple: same f32 products
    and the same tap order, so the only differences are the few-ulp libm differences of the taps."""
    from conftest import make_model
    cfg, w, m = make_model("qwen3_asr", "tiny", max_ctx=64, max_frames=100)
    try:
        rng = np.random.default_rng(orig + n)
        x = (0.3 * rng.standard_normal(n)).astype(np.float32)
        got = m.resample(x, orig, new)
        want = OA.resample_simple(x[None], orig, new)[0]
        assert got.shape == want.shape
        if n:
            assert np.abs(got - want).max() <= 2e-6
    finally:
        m.close()
