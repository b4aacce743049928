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
            assert np.array_equal(pv, want_pv)
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


# ---- video half of Qwen3VLProcessor (qwen3vl/processor.rs:253-307, 404-437, 447-571; utils/video_utils.rs:9-59) ----
def test_video_smart_resize_matches_the_oracle_and_known_answers():
    # 16 frames of 1080p: 16 * 1088 * 1920 > 25165824 -> beta = sqrt(16*1080*1920 / 25165824), floors to multiples of lcm(32, 16) = 32
    assert P.video_smart_resize(16, 1080, 1920) == OV.video_smart_resize(16, 1080, 1920, 2, 32, 4096, 25165824, 16) == (928, 1664)
    assert P.video_smart_resize(4, 360, 640) == (352, 640)                      # inside the budget: plain rounding to the factor
    assert P.video_smart_resize(4, 360, 640, factor=28, video_ratio=16) == OV.video_smart_resize(4, 360, 640, 2, 28, 4096, 25165824, 16)  # lcm(28, 16) = 112
    assert P.video_smart_resize(4, 360, 640, factor=28, video_ratio=0) == OV.video_smart_resize(4, 360, 640, 2, 28, 4096, 25165824, None) == (364, 644)
    rng = np.random.default_rng(3)
    for _ in range(300):
        t, h, w = int(rng.integers(2, 769)), int(rng.integers(32, 2200)), int(rng.integers(32, 4000))
        if max(h, w) // min(h, w) > 200:
            continue
        mn, mx = int(rng.choice([4096, 262144, 1 << 22])), int(rng.choice([1 << 20, 25165824, 1 << 28]))
        assert P.video_smart_resize(t, h, w, 2, 32, mn, mx, 16) == OV.video_smart_resize(t, h, w, 2, 32, mn, mx, 16), (t, h, w, mn, mx)
    with pytest.raises(P.ProcessorError, match="must be larger than temporal_factor"):
        P.video_smart_resize(1, 360, 640)
    with pytest.raises(P.ProcessorError, match="must be larger than factor"):
        P.video_smart_resize(4, 16, 640)
    with pytest.raises(P.ProcessorError, match="aspect ratio"):
        P.video_smart_resize(4, 32, 32 * 300)


def test_video_frame_sampling_and_timestamps():
    # 10 s at 30 fps, 2 samples per second: 20 frames, every 15th
    nf, idx = P.video_sample_frames(300, 30)
    assert nf == 20 and idx.tolist() == list(range(0, 300, 15))
    # 1 s at 25 fps: round(2) = 2 < min_frames 4 -> 4 frames, interval round(25 / 4) = 6 -> 5 frames are kept (the reference's own gap)
    nf, idx = P.video_sample_frames(25, 25)
    assert nf == 4 and idx.tolist() == [0, 6, 12, 18, 24]
    # NTSC 30000/1001, an hour: clamped to max_frames 768
    nf, idx = P.video_sample_frames(107892, 30000, 1001)
    assert nf == 768 and idx[1] == round(107892 / 768)
    rng = np.random.default_rng(4)
    for _ in range(200):
        total, num, den = int(rng.integers(1, 20000)), int(rng.choice([24, 25, 30, 60, 24000, 30000])), 1
        if num > 1000:
            den = 1001
        fps, mn, mx = int(rng.choice([1, 2, 4])), int(rng.choice([1, 4, 16])), int(rng.choice([8, 64, 768]))
        nf, idx = P.video_sample_frames(total, num, den, fps, mn, mx)
        want_nf, want_idx = OV.video_sample_frames(total, num, den, fps, mn, mx)
        assert nf == want_nf and idx.tolist() == want_idx, (total, num, den, fps, mn, mx)
    # calculate_timestamps: mean time of the first and last frame of each pair; an odd count repeats the last frame
    assert P.video_timestamps([0, 15, 30, 45], 30.0).tolist() == [0.25, 1.25]
    assert P.video_timestamps([0, 15, 30, 45, 60], 30.0).tolist() == [0.25, 1.25, 2.0]
    for n in (1, 2, 7, 20, 33):
        idx = np.sort(rng.choice(5000, n, replace=False))
        for fps in (23.976, 25.0, 29.97, 60.0):
            got = P.video_timestamps(idx, fps)
            want = np.asarray(OV.calculate_timestamps(idx.tolist(), fps, 2), np.float32)
            assert np.array_equal(got, want), (n, fps)
    # "<{:.1} seconds>": exact decimal of the f32, ties to even
    assert [P.format_timestamp(t) for t in (0.25, 0.75, 1.25, 2.0, 12.349999, 0.05)] == \
           ["<0.2 seconds>", "<0.8 seconds>", "<1.2 seconds>", "<2.0 seconds>", "<12.3 seconds>", "<0.1 seconds>"]
    for t in rng.random(100).astype(np.float32) * 100:
        assert P.format_timestamp(t) == OV.format_timestamp(t)


def test_video_placeholder_expansion_matches_the_string_edit():
    """The token-id expansion against the reference's own string editing (restated verbatim in the oracle) through a toy tokenizer
    in which every special token and every character is one id."""
    VS, VP, VE = "<|vision_start|>", "<|video_pad|>", "<|vision_end|>"
    special = {VS: 1, VP: 2, VE: 3}

    def tok(text):
        ids = []
        while text:
            for s, i in special.items():
                if text.startswith(s):
                    ids.append(i); text = text[len(s):]; break
            else:
                ids.append(1000 + ord(text[0])); text = text[1:]
        return ids

    cases = [
        ("hi " + VS + VP + VE + " what happens?", [[2, 4, 6]]),
        ("a" + VP + "b", [[3, 2, 2]]),                                             # lone pad: no start/end tokens around the expansion
        (VP + " then " + VS + VP + VE, [[1, 2, 2], [2, 2, 4]]),                      # the triple is replaced first, by video 0
        (VS + VP + VE + VS + VP + VE + "x", [[1, 4, 4], [2, 2, 2]]),
        ("no video here", []),
    ]
    for text, grids in cases:
        stamps = [OV.calculate_timestamps(list(range(0, 2 * g[0] * 7, 7)), 30.0, 2) for g in grids]
        want = tok(OV.expand_video_placeholders_text(text, grids, stamps))
        runs = [tok(OV.format_timestamp(t)) for st in stamps for t in st]
        got = P.expand_video_placeholders(tok(text), grids, runs, 2, 1, 3)
        assert got.tolist() == want, text
    with pytest.raises(P.ProcessorError, match="more <.video_pad.> placeholders"):
        P.expand_video_placeholders(tok(VP + VP), [[1, 2, 2]], [tok("<0.0 seconds>")], 2, 1, 3)
    with pytest.raises(P.ProcessorError, match="fewer timestamp"):
        P.expand_video_placeholders(tok(VP), [[2, 2, 2]], [tok("<0.0 seconds>")], 2, 1, 3)


def test_oracle_process_video_known_answers():
    """process_videos restated: T frames pad to a multiple of 2 with the last frame; a clip of identical frames gives the image rows."""
    img = synth.synth_image(64, 96, 5)
    pv_img, grid_img = OV.process_image(img, min_pixels=1024)          # no resize: 64 x 96 already fits
    pv, grid = OV.process_video(np.stack([img, img]))
    assert grid.tolist() == [[1, 4, 6]] and np.array_equal(pv, pv_img)
    frames = np.stack([synth.synth_image(64, 96, s) for s in (1, 2, 3)])
    pv3, grid3 = OV.process_video(frames)
    assert grid3.tolist() == [[2, 4, 6]] and pv3.shape == (48, 1536)
    pv4, _ = OV.process_video(np.concatenate([frames, frames[2:3]]))
    assert np.array_equal(pv3, pv4)
    # feature order (c, frame in group, py, px): patch 0, channel 0, second frame, pixel (0, 0) = frame 1's pixel
    assert pv3[0, 256] == np.float32((np.float32(frames[1, 0, 0, 0]) * np.float32(1 / 255) - np.float32(0.5)) / np.float32(0.5))


@pytest.mark.gpu
@pytest.mark.parametrize("n_frames", [2, 5, 8])
def test_gpu_video_pipeline_matches_the_oracle(n_frames):
    """The whole video request the way Qwen3VLProcessor::process_info builds it -- sampling, video_smart_resize, process_videos on the GPU,
    timestamps, <|video_pad|> expansion -- then generate_generic on the expanded prompt against the oracle on the oracle's own tensors."""
    from conftest import make_model, make_oracle
    from oracle.generate import GenerationContext, generate_generic
    cfg, w, m = make_model("qwen3vl", "tiny", max_ctx=1024, max_patches=1024)
    o = make_oracle("qwen3vl", cfg, w)
    vc = cfg["vision_config"]
    # a 3 s clip at 25 fps of 70 x 100 frames; the processor keeps `n_frames` of them (min_frames = max_frames = n_frames)
    nf, idx = P.video_sample_frames(75, 25, 1, 2, n_frames, n_frames)
    rh, rw = P.video_smart_resize(nf, 70, 100, vc["temporal_patch_size"], vc["patch_size"] * vc["spatial_merge_size"], 4096, 25165824, 16)
    assert (rh, rw) == OV.video_smart_resize(nf, 70, 100, 2, 32, 4096, 25165824, 16) == (64, 96)
    frames = np.stack([synth.synth_image(rh, rw, 100 + int(i)) for i in idx])            # what the scaler would hand over, RGB24
    pvv, vgrid = m.video_preprocess(frames)
    want_pvv, want_grid = OV.process_video(frames)
    assert vgrid.tolist() == want_grid.tolist() == [[(len(idx) + 1) // 2, 2 * 2, 3 * 2]]
    assert np.array_equal(pvv, want_pvv)                                                  # one multiply, one subtract, one divide per pixel: bit-exact
    stamps = P.video_timestamps(idx, 25.0, vc["spatial_merge_size"])                      # the reference passes merge_size here (processor.rs:411-415)
    assert np.array_equal(stamps, np.asarray(OV.calculate_timestamps(idx.tolist(), 25.0, 2), np.float32))
    toy_tok = lambda s: [300 + (ord(c) % 64) for c in s]                                  # stand-in for the tokenizer (out of scope, stays with aha)
    runs = [toy_tok(P.format_timestamp(t)) for t in stamps]
    prompt = np.concatenate([synth.synth_text_ids(4, 1000, 1), [cfg["vision_start_token_id"], cfg["video_token_id"], cfg["vision_end_token_id"]],
                             synth.synth_text_ids(7, 1000, 2)]).astype(np.uint32)
    ids = P.expand_video_placeholders(prompt, vgrid, runs, cfg["video_token_id"], cfg["vision_start_token_id"], cfg["vision_end_token_id"],
                                      vc["spatial_merge_size"])
    assert (ids == cfg["video_token_id"]).sum() == int(np.prod(vgrid)) // 4
    data = [None, None, pvv, vgrid, None]
    got = m.forward_initial(ids, 0, data)[0, 0]
    want = o.forward_initial(ids.reshape(1, -1), 0, [None, None, want_pvv, want_grid, None])[0, 0]
    assert np.abs(got - want).max() <= 1e-3
    m.clear_cache(); o.clear_cache()
    ctx = GenerationContext(temperature=0.0, initial_seq_len=len(ids), max_tokens=6)
    want_toks, _, _ = generate_generic(o, ids.reshape(1, -1), [None, None, want_pvv, want_grid, None], ctx)
    toks, usage = m.generate(ids, data, max_tokens=6)
    assert toks == want_toks and usage["vision_secs"] > 0
    with pytest.raises(Exception, match="multiple of patch_size"):
        m.video_preprocess(np.zeros((2, 70, 100, 3), np.uint8))
    m.close()


def test_video_placeholder_expansion_property():
    """Random prompts (text, lone <|video_pad|>, full <|vision_start|><|video_pad|><|vision_end|> triples, stray start / end tokens) and random
    grids: the id-level expansion equals the tokenised result of the reference's string edit, every time."""
    from hypothesis import assume, given, settings, strategies as st
    VS, VP, VE = "<|vision_start|>", "<|video_pad|>", "<|vision_end|>"
    special = {VS: 1, VP: 2, VE: 3}

    def tok(text):
        ids = []
        while text:
            for s, i in special.items():
                if text.startswith(s):
                    ids.append(i); text = text[len(s):]; break
            else:
                ids.append(1000 + ord(text[0])); text = text[1:]
        return ids

    piece = st.sampled_from(["a", "bc ", VP, VS + VP + VE, VS, VE, " ", VS + VP, VP + VE])

    @settings(max_examples=150, deadline=None)
    @given(st.lists(piece, max_size=9), st.lists(st.tuples(st.integers(1, 3), st.sampled_from([2, 4]), st.sampled_from([2, 4, 6])), min_size=6, max_size=6))
    def check(pieces, grids):
        text = "".join(pieces)
        n_videos = text.count(VP)
        assume(n_videos <= len(grids))
        grids_used = [list(g) for g in grids[:max(n_videos, 1)]]
        stamps = [OV.calculate_timestamps(list(range(0, 2 * g[0] * 5, 5)), 25.0, 2) for g in grids_used]
        want = tok(OV.expand_video_placeholders_text(text, grids_used, stamps)) if n_videos else tok(text)
        runs = [tok(OV.format_timestamp(t)) for stm in stamps for t in stm]
        got = P.expand_video_placeholders(tok(text), grids_used, runs, 2, 1, 3)
        assert got.tolist() == want, text

    check()
