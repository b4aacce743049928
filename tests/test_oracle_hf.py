"""Cross-check of the oracle against HF transformers where the reference has no quirk (text decoder, ViT,
pos-embed interpolation, 2-D RoPE, deepstack, M-RoPE, get_rope_index, slaney mel bank, the audio tower with the
reference's two deviations from HF -- tanh-GELU in the conv stem, no attention windows -- neutralised).  CPU only."""
import numpy as np
import pytest

from aha_b200 import synth

torch = pytest.importorskip("torch")
transformers = pytest.importorskip("transformers")


@pytest.mark.parametrize("preset", ["tiny", "tiny-g1", "tiny-g4"])   # GQA groups 2, 1 (MHA) and 4
def test_qwen3_matches_hf(preset):
    from transformers import Qwen3Config, Qwen3ForCausalLM
    from oracle.qwen3 import Qwen3Model
    cfg = synth.get_config("qwen3", preset)
    w = synth.make_weights("qwen3", cfg, 0)
    hc = Qwen3Config(**cfg, max_position_embeddings=4096)
    hc._attn_implementation = "eager"
    hf = Qwen3ForCausalLM(hc).float().eval()
    sd = {k: torch.from_numpy(v.astype(np.float32)) for k, v in w.items()}
    sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
    hf.load_state_dict(sd, strict=False)
    ids = synth.synth_text_ids(12, 1000, 5).astype(np.int64)
    m = Qwen3Model(cfg, w)
    got = m.forward_step(ids.reshape(1, -1), 0)[0, 0]
    with torch.no_grad():
        want = hf(torch.from_numpy(ids)[None]).logits[0, -1].numpy()
    assert np.abs(got - want).max() < 1e-5
    got2 = m.forward_step(np.array([[7]]), 12)[0, 0]          # decode step against the oracle's cat-cache
    with torch.no_grad():
        want2 = hf(torch.from_numpy(np.concatenate([ids, [7]]))[None]).logits[0, -1].numpy()
    assert np.abs(got2 - want2).max() < 1e-5


def test_qwen3vl_matches_hf():
    from transformers import Qwen3VLConfig, Qwen3VLForConditionalGeneration
    from oracle.qwen3vl import Qwen3VLModel, get_rope_index, process_image
    cfg = synth.get_config("qwen3vl", "tiny")
    w = synth.make_weights("qwen3vl", cfg, 0)
    hc = Qwen3VLConfig(text_config=dict(cfg["text_config"], max_position_embeddings=4096), vision_config=cfg["vision_config"],
                       image_token_id=cfg["image_token_id"], video_token_id=cfg["video_token_id"],
                       vision_start_token_id=cfg["vision_start_token_id"], vision_end_token_id=cfg["vision_end_token_id"],
                       tie_word_embeddings=True)
    for c in (hc, hc.vision_config, hc.text_config):
        c._attn_implementation = "eager"
    hf = Qwen3VLForConditionalGeneration(hc).float().eval()
    sd = {k: torch.from_numpy(v.astype(np.float32)) for k, v in w.items()}
    sd["lm_head.weight"] = sd["model.language_model.embed_tokens.weight"]
    hf.load_state_dict(sd, strict=False)
    pv, grid = process_image(synth.synth_image(256, 320, 1))
    ids = np.concatenate([synth.synth_text_ids(3, 1000, 9), synth.vl_prompt_ids(cfg, grid, 9)]).astype(np.int64)
    pos, delta = get_rope_index(ids, grid, cfg)
    mm = torch.from_numpy((ids == cfg["image_token_id"]).astype(np.int64))[None]
    tg = torch.from_numpy(grid.astype(np.int64))
    with torch.no_grad():
        try:
            out = hf(input_ids=torch.from_numpy(ids)[None], pixel_values=torch.from_numpy(pv), image_grid_thw=tg,
                     mm_token_type_ids=mm).logits[0, -1].numpy()
        except TypeError:
            out = hf(input_ids=torch.from_numpy(ids)[None], pixel_values=torch.from_numpy(pv), image_grid_thw=tg).logits[0, -1].numpy()
    assert int(hf.model.rope_deltas.reshape(-1)[0]) == delta
    m = Qwen3VLModel(cfg, w)
    got = m.forward_initial(ids.reshape(1, -1), 0, [pv, grid, None, None, None])[0, 0]
    assert np.abs(got - out).max() < 1e-5


def test_qwen3vl_video_branch_matches_hf():
    """Image + video in one prompt: the video tensors through the same tower, scattered at <|video_pad|>, joint deepstack injection, M-RoPE
    with every temporal step of the video as its own (1, h, w) grid (model.rs:907-925, 1169-1225) -- against HF's implementation."""
    from transformers import Qwen3VLConfig, Qwen3VLForConditionalGeneration
    from oracle.qwen3vl import Qwen3VLModel, get_rope_index, img_transform, process_image, process_vision_tensor
    cfg = synth.get_config("qwen3vl", "tiny")
    w = synth.make_weights("qwen3vl", cfg, 0)
    hc = Qwen3VLConfig(text_config=dict(cfg["text_config"], max_position_embeddings=4096), vision_config=cfg["vision_config"],
                       image_token_id=cfg["image_token_id"], video_token_id=cfg["video_token_id"],
                       vision_start_token_id=cfg["vision_start_token_id"], vision_end_token_id=cfg["vision_end_token_id"],
                       tie_word_embeddings=True)
    for c in (hc, hc.vision_config, hc.text_config):
        c._attn_implementation = "eager"
    hf = Qwen3VLForConditionalGeneration(hc).float().eval()
    sd = {k: torch.from_numpy(v.astype(np.float32)) for k, v in w.items()}
    sd["lm_head.weight"] = sd["model.language_model.embed_tokens.weight"]
    hf.load_state_dict(sd, strict=False)
    frames = np.stack([img_transform(synth.synth_image(64, 96, 11 + i), (0.5,) * 3, (0.5,) * 3) for i in range(4)])
    pvv, vgrid = process_vision_tensor(frames)                        # 4 frames -> video_grid_thw [[2, 4, 6]]
    pv, grid = process_image(synth.synth_image(128, 96, 1))
    ids = np.concatenate([synth.synth_text_ids(3, 1000, 9), synth.vl_prompt_ids(cfg, grid, 0), synth.vl_video_prompt_ids(cfg, vgrid),
                          synth.synth_text_ids(6, 1000, 4)]).astype(np.int64)
    _, delta = get_rope_index(ids, grid, cfg, vgrid)
    mm = torch.from_numpy((ids == cfg["image_token_id"]).astype(np.int64) + 2 * (ids == cfg["video_token_id"]).astype(np.int64))[None]
    kw = dict(input_ids=torch.from_numpy(ids)[None], pixel_values=torch.from_numpy(pv), image_grid_thw=torch.from_numpy(grid.astype(np.int64)),
              pixel_values_videos=torch.from_numpy(pvv), video_grid_thw=torch.from_numpy(vgrid.astype(np.int64)))
    with torch.no_grad():
        try:
            out = hf(**kw, mm_token_type_ids=mm).logits[0, -1].numpy()
        except TypeError:
            out = hf(**kw).logits[0, -1].numpy()
    assert int(hf.model.rope_deltas.reshape(-1)[0]) == delta
    m = Qwen3VLModel(cfg, w)
    got = m.forward_initial(ids.reshape(1, -1), 0, [pv, grid, pvv, vgrid, None])[0, 0]
    assert np.abs(got - out).max() < 1e-5


def test_mel_filter_bank_matches_hf():
    from transformers.audio_utils import mel_filter_bank as hfmel
    from oracle.audio import mel_filter_bank
    h = hfmel(201, 128, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney")
    assert np.abs(h - mel_filter_bank(201, 128, 0.0, 8000.0, 16000)).max() < 1e-6


def test_audio_tower_matches_hf_omni_encoder():
    """The ASR audio tower against HF's Qwen3-Omni audio encoder (the architecture it was derived from).  The two places
    where the reference differs are neutralised: the conv-stem GELU is switched to erf in the oracle
    (qwen3_asr/model.rs:200-202 uses the tanh form; covered by test_activations_match_definitions), HF's attention
    window is made larger than the sequence (the reference attends over all chunks, model.rs:218-220), and the oracle's
    sinusoid table gets HF's timescales: the reference builds it from the RoPE helper, 10000^(-i/(d/2))
    (sinusoidal_pe.rs:13, rope.rs:7-13), where HF/Whisper use 10000^(-i/(d/2-1)) -- a third deviation this test found;
    the oracle and the CUDA path keep the reference's form."""
    try:
        from transformers.models.qwen3_omni_moe.configuration_qwen3_omni_moe import Qwen3OmniMoeAudioEncoderConfig
        from transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe import Qwen3OmniMoeAudioEncoder
    except ImportError:
        pytest.skip("transformers build without qwen3_omni_moe")
    from oracle import nn as onn
    from oracle.audio import get_feat_extract_output_lengths
    from oracle.qwen3_asr import Qwen3ASRAudioEncoder
    cfg = synth.get_config("qwen3_asr", "tiny")
    ac = cfg["thinker_config"]["audio_config"]
    w = synth.make_weights("qwen3_asr", cfg, 0)
    hc = Qwen3OmniMoeAudioEncoderConfig(**dict(ac, n_window_infer=100 * 1000))
    hc._attn_implementation = "eager"
    hf = Qwen3OmniMoeAudioEncoder(hc).float().eval()
    pre = "thinker.audio_tower."
    sd = {k[len(pre):]: torch.from_numpy(v.astype(np.float32)) for k, v in w.items() if k.startswith(pre)}
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected and all("positional_embedding" in k for k in missing), (missing, unexpected)
    T = 250                                                   # chunks of 100, 100, 50 frames -> 13 + 13 + 7 tokens
    rng = np.random.default_rng(3)
    mel = rng.standard_normal((ac["num_mel_bins"], T)).astype(np.float32)
    enc = Qwen3ASRAudioEncoder(ac, w)
    enc.conv_act = onn.gelu_erf
    half = ac["d_model"] // 2
    ref_form = enc.pe.inv_freq.copy()
    enc.pe.inv_freq = np.exp(-np.log(10000.0) / (half - 1) * np.arange(half, dtype=np.float32))[None].astype(np.float32)
    assert np.abs(ref_form - np.float32(10000.0) ** (-np.arange(half, dtype=np.float32) / np.float32(half))).max() < 1e-6   # sinusoidal_pe.rs:13
    got = enc.forward(mel)
    with torch.no_grad():
        want = hf(torch.from_numpy(mel), feature_lens=torch.tensor([T]),
                  aftercnn_lens=torch.tensor([get_feat_extract_output_lengths(T)])).last_hidden_state.numpy()
    assert got.shape == want.shape == (33, ac["output_dim"])
    assert np.abs(got - want).max() < 2e-5


def test_log_mel_matches_hf_whisper_frontend():
    """The whole log-mel frontend (framing, power spectrum, slaney bank, log10, -8 dB clamp, (x+4)/4) against HF's
    WhisperFeatureExtractor on 30 s of audio.  Neutralised: the Hann window (the reference's is symmetric,
    audio_utils.rs:1071-1082; HF's periodic).  Excluded: the last frame, the only one that reaches into the right
    reflect pad, which the reference cuts from the already left-padded tensor (tensor_utils.rs:525-549)."""
    from transformers import WhisperFeatureExtractor as HFExtractor
    from oracle.audio import WhisperFeatureExtractor
    wav = synth.synth_audio(30.0, 16000, 2)
    want = HFExtractor(feature_size=128)(wav, sampling_rate=16000, return_tensors="np", padding="max_length")["input_features"][0]
    fe = WhisperFeatureExtractor()
    symmetric = fe.extract_fbank_features(wav[None])[0]
    fe.window = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(400) / 400)).astype(np.float32)
    got = fe.extract_fbank_features(wav[None])[0]
    assert got.shape == want.shape == (128, 3000)
    assert np.abs(got[:, :-1] - want[:, :-1]).max() < 5e-4        # f32 power spectrum vs HF's f64, in the log domain
    assert np.abs(got[:, -1] - want[:, -1]).max() > 1e-2          # the reflect-pad quirk is real ...
    assert np.abs(symmetric[:, :-1] - want[:, :-1]).max() > 1e-2  # ... and so is the window's


def test_image_patchify_matches_hf_processor():
    """img_transform + the 9-D patchify permute (qwen3vl/processor.rs:151-251) against HF's Qwen2VLImageProcessor for an
    image that needs no resize (the CatmullRom resize itself is out of scope)."""
    from transformers import Qwen2VLImageProcessor
    from oracle.qwen3vl import process_image
    for h, w in ((256, 320), (64, 1024)):
        img = synth.synth_image(h, w, 1)
        pv, grid = process_image(img)
        ip = Qwen2VLImageProcessor(patch_size=16, temporal_patch_size=2, merge_size=2, image_mean=[0.5] * 3, image_std=[0.5] * 3,
                                   min_pixels=65536, max_pixels=16777216)
        out = ip(images=[img], return_tensors="np")
        assert (np.asarray(out["image_grid_thw"]) == grid).all()
        assert np.abs(out["pixel_values"] - pv).max() < 1e-6


def test_get_rope_index_matches_hf_on_interleaved_images():
    """M-RoPE position ids (qwen3vl/model.rs:901-1133) for prompts with several images of different grids separated by
    text, against HF's Qwen3VLModel.get_rope_index -- every position of all three rows, and rope_deltas."""
    from transformers import Qwen3VLConfig, Qwen3VLForConditionalGeneration
    from oracle.qwen3vl import get_rope_index
    cfg = synth.get_config("qwen3vl", "tiny")
    hc = Qwen3VLConfig(text_config=dict(cfg["text_config"], max_position_embeddings=4096), vision_config=cfg["vision_config"],
                       image_token_id=cfg["image_token_id"], video_token_id=cfg["video_token_id"],
                       vision_start_token_id=cfg["vision_start_token_id"], vision_end_token_id=cfg["vision_end_token_id"],
                       tie_word_embeddings=True)
    hf = Qwen3VLForConditionalGeneration(hc).eval()
    cases = [
        ([[1, 4, 6]], [5, 7]),                          # text, image, text
        ([[1, 4, 6], [1, 8, 2]], [3, 4, 9]),            # two images, text on every side
        ([[1, 2, 2], [1, 6, 10], [1, 4, 4]], [0, 1, 0, 6]),   # leading image, adjacent images, trailing text
        ([[1, 68, 120]], [0, 512]),                     # the north-star prompt shape: 2040 image tokens + 512 ids
    ]
    for grids, texts in cases:
        ids = []
        for k, n in enumerate(texts):
            ids += synth.synth_text_ids(n, 1000, 20 + k).tolist()
            if k < len(grids):
                ids += synth.vl_prompt_ids(cfg, [grids[k]], 0).tolist()
        ids = np.asarray(ids, dtype=np.int64)
        grid = np.asarray(grids, dtype=np.int64)
        pos, delta = get_rope_index(ids, grid, cfg)
        mm = torch.from_numpy((ids == cfg["image_token_id"]).astype(np.int32))[None]
        want_pos, want_delta = hf.model.get_rope_index(torch.from_numpy(ids)[None], mm, image_grid_thw=torch.from_numpy(grid))
        assert (pos == want_pos.numpy()).all(), (grids, texts)
        assert delta == int(want_delta.reshape(-1)[0])
    pos, delta = get_rope_index(ids, grid, cfg)
    assert delta == -1980                               # max(34, 60) - 2040


def test_video_processor_restatement_matches_hf():
    """The video half of the processor restated from the reference (qwen3vl/processor.rs:253-280, utils/video_utils.rs:9-59) against HF's
    Qwen3VLVideoProcessor: the patch tensor of a clip (odd frame counts pad with the last frame) and the resize target.  The reference multiplies
    frames * height * width in u32: wherever that product fits, its answer is HF's; where it wraps, the restatement follows the wrap (documented)."""
    import torch
    from transformers.models.qwen3_vl import video_processing_qwen3_vl as V
    from aha_b200 import synth
    from oracle import qwen3vl as OV
    vp = V.Qwen3VLVideoProcessor(patch_size=16, temporal_patch_size=2, merge_size=2, image_mean=[0.5] * 3, image_std=[0.5] * 3,
                                 do_resize=False, do_sample_frames=False)
    for T in (2, 3, 5, 8):
        frames = np.stack([synth.synth_image(64, 96, 10 + i) for i in range(T)])
        out = vp(videos=[torch.from_numpy(frames).permute(0, 3, 1, 2)], return_tensors="pt")
        pv, grid = OV.process_video(frames)
        assert out["video_grid_thw"].numpy().tolist() == grid.tolist() == [[(T + 1) // 2, 4, 6]]
        assert np.abs(out["pixel_values_videos"].numpy() - pv).max() <= 2e-7
    rng = np.random.default_rng(0)
    n = 0
    for _ in range(1500):
        t, h, w = int(rng.integers(1, 400)) * 2, int(rng.integers(32, 2200)), int(rng.integers(32, 4000))
        if max(h, w) / min(h, w) > 200 or (h / 32) % 1 == 0.5 or (w / 32) % 1 == 0.5 or t * h * w >= 1 << 32:
            continue                                    # f32 round-half-away vs Python's round-half-even on exact halves; the u32 wrap
        mn, mx = int(rng.choice([4096, 262144])), int(rng.choice([1 << 20, 25165824]))
        assert OV.video_smart_resize(t, h, w, 2, 32, mn, mx, None) == V.smart_resize(t, h, w, 2, 32, mn, mx), (t, h, w, mn, mx)
        n += 1
    assert n > 1000
    # the wrap itself: 628 frames of 2068 x 3853 is 5.0e9 pixels; the reference sees 5.0e9 mod 2^32 = 7.1e8 and scales less than HF would
    assert OV.video_smart_resize(628, 2068, 3853, 2, 32, 262144, 25165824, None) == (384, 704)
    assert V.smart_resize(628, 2068, 3853, 2, 32, 262144, 25165824) == (128, 256)


def test_catmullrom_resize_restatement_matches_pillow_bicubic():
    """`image`'s resize(CatmullRom) restated in oracle/qwen3vl.py (qwen3vl/processor.rs:167) against an independent implementation of the same
    filter: Pillow's BICUBIC is the a = -0.5 cubic with the same centre alignment and the same support scaling when shrinking.  Pillow keeps an 8-bit
    intermediate between its two passes (the crate keeps f32), so the comparison is on images whose intermediate cannot clip -- smooth ones at any
    scale, and noise when shrinking (averaging keeps it inside 0..255) -- and allows the LSB that Pillow's fixed-point coefficients cost (two on a handful of pixels when ~10 taps meet noise)."""
    PIL = pytest.importorskip("PIL.Image")
    from aha_b200 import synth
    from oracle.qwen3vl import resize_exact_catmullrom
    yy, xx = np.mgrid[0:300, 0:400]
    smooth = np.stack([(127 + 100 * np.sin(xx / 17.0) * np.cos(yy / 23.0)).astype(np.uint8),
                       (127 + 90 * np.cos(xx / 29.0 + yy / 41.0)).astype(np.uint8),
                       ((xx + yy) % 256 // 2 + 60).astype(np.uint8)], -1)
    cases = [(smooth, nh, nw) for nh, nw in ((128, 160), (600, 800), (300, 400), (77, 391), (512, 96))]
    cases += [(synth.synth_image(h, w, h + w), nh, nw) for h, w, nh, nw in ((300, 200, 96, 64), (480, 640, 96, 160), (1080, 1920, 352, 640))]
    for img, nh, nw in cases:
        got = resize_exact_catmullrom(img, nh, nw)
        want = np.asarray(PIL.fromarray(img).resize((nw, nh), PIL.BICUBIC))
        d = np.abs(got.astype(int) - want.astype(int))
        assert d.max() <= 2 and (d > 1).mean() < 1e-3, (img.shape, nh, nw, int(d.max()), float((d > 1).mean()))   # 2 only with ~10 taps per pixel
        assert (d > 0).mean() < 0.35


def test_img_smart_resize_restatement_matches_hf():
    """img_smart_resize (img_utils.rs:297-331) against HF's Qwen2-VL smart_resize: equal wherever no side rounds to zero (the reference clamps a
    side to the factor BEFORE it compares the pixel count with the budget, HF after) and no side sits on an exact half (f32 round-half-away vs
    Python's round-half-even)."""
    from transformers.models.qwen2_vl.image_processing_qwen2_vl import smart_resize
    from oracle.qwen3vl import img_smart_resize
    rng = np.random.default_rng(1)
    n = 0
    for _ in range(3000):
        h, w = int(rng.integers(16, 6000)), int(rng.integers(16, 6000))
        if max(h, w) / min(h, w) > 200 or (h / 32) % 1 == 0.5 or (w / 32) % 1 == 0.5:
            continue
        mn, mx = int(rng.choice([3136, 65536, 262144])), int(rng.choice([1048576, 16777216]))
        assert img_smart_resize(h, w, 32, mn, mx) == smart_resize(h, w, 32, mn, mx), (h, w, mn, mx)
        n += 1
    assert n > 2500
    assert img_smart_resize(15, 1425, 32, 3136, 1048576) == (32, 1440) and smart_resize(15, 1425, 32, 3136, 1048576) == (32, 576)   # the clamp-first quirk


def test_sinc_resampler_restatement_matches_torchaudio():
    """resample_simple (audio_utils.rs:66-255) is torchaudio's sinc_interp_hann resampler (lowpass_filter_width 6, rolloff 0.99) in f32: the restatement
    against torchaudio.functional.resample itself -- a few 1e-7 for small filter banks, a few 1e-5 where 477 f32 taps meet torchaudio's f64-built kernel."""
    torch = pytest.importorskip("torch")
    torchaudio = pytest.importorskip("torchaudio")
    from oracle.audio import resample_simple
    rng = np.random.default_rng(0)
    for orig, new, n, tol in ((48000, 16000, 48000, 1e-6), (44100, 16000, 30011, 5e-5), (8000, 16000, 4001, 1e-6), (22050, 16000, 22050, 1e-4),
                              (24000, 16000, 7, 1e-6), (16000, 24000, 1001, 1e-6)):
        x = (0.3 * rng.standard_normal(n)).astype(np.float32)
        got = resample_simple(x[None], orig, new)[0]
        want = torchaudio.functional.resample(torch.from_numpy(x)[None], orig, new, lowpass_filter_width=6, rolloff=0.99,
                                              resampling_method="sinc_interp_hann")[0].numpy()
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= tol, (orig, new, float(np.abs(got - want).max()))
