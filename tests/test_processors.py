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
