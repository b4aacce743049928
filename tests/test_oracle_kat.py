"""Known answers derivable from the reference source alone pin the oracle,
since aha's own tests hold no golden vectors."""
import numpy as np

from aha_b200 import synth
from oracle import audio, nn, qwen3, qwen3vl, rope


def test_img_smart_resize_1080p():  # img_utils.rs:297-331
    assert qwen3vl.img_smart_resize(1080, 1920, 32, 65536, 16777216) == (1088, 1920)
    assert qwen3vl.img_smart_resize(2048, 2048, 32, 65536, 16777216) == (2048, 2048)
    assert qwen3vl.img_smart_resize(64, 96, 32, 65536, 16777216) == (224, 320)  # below min_pixels: scaled up


def test_patchify_shapes_1080p():  # processor.rs:186-226
    img = np.zeros((1088, 1920, 3), np.uint8)
    pv, grid = qwen3vl.process_image(img)
    assert pv.shape == (8160, 1536) and grid.tolist() == [[1, 68, 120]]
    assert 68 * 120 // 4 == 2040


def test_patchify_order_is_merge_block_major():
    h, w = 256, 320
    img = np.zeros((h, w, 3), np.uint8)
    img[16:32, 0:16, 0] = 255      # patch (row 1, col 0)
    pv, grid = qwen3vl.process_image(img)
    lit = np.nonzero(pv[:, 0] > 0)[0].tolist()
    assert lit == [2]              # block (0,0): patches (0,0),(0,1),(1,0),(1,1) -> index 2
    # feature order (c, t, ph, pw): channel 0 both frames lit, channel 1 dark
    assert np.all(pv[2, :512] == 1.0) and np.all(pv[2, 512:] == -1.0)


def test_rope_delta_single_1080p_image():  # model.rs:982-1071
    cfg = synth.get_config("qwen3vl", "vl2")
    grid = np.array([[1, 68, 120]])
    for n_text in (0, 7, 512):
        ids = synth.vl_prompt_ids(cfg, grid, n_text)
        pos, delta = qwen3vl.get_rope_index(ids, grid, cfg)
        assert delta == max(34, 60) - 2040 == -1980
        assert pos.shape == (3, 1, len(ids))
        assert pos[:, 0, 0].tolist() == [0, 0, 0]          # <|vision_start|>
        assert pos[:, 0, 1].tolist() == [1, 1, 1]          # first image token (t, h, w) = base
        assert pos[:, 0, 2040].tolist() == [1, 34, 60]     # last image token: base + (0, 33, 59)
        assert pos[:, 0, 2041].tolist() == [61, 61, 61]    # <|vision_end|> starts at max + 1


def test_feat_extract_output_lengths():  # qwen3_asr/processor.rs:187-195
    assert [audio.get_feat_extract_output_lengths(n) for n in (1, 100, 250, 3000)] == [1, 13, 33, 390]


def test_mel_frame_count_30s():  # feature_extraction_whisper.rs:98-105
    fe = audio.WhisperFeatureExtractor()
    mel = fe.call(synth.synth_audio(3.0)[None], 16000)
    assert mel.shape == (1, 128, 300)
    assert 480000 // 160 == 3000


def test_hann_is_symmetric_not_periodic():  # audio_utils.rs:1071-1082
    w = audio.create_hann_window(400)
    k = np.arange(400)
    assert w[0] == 0.0 and w[399] == 0.0
    np.testing.assert_allclose(w, 0.5 - 0.5 * np.cos(2 * np.pi * k / 399.0), atol=1e-6)


def test_reflect_pad_quirk():  # tensor_utils.rs:525-549: right pad cut from the left-padded tensor
    x = np.arange(10, dtype=np.float32)[None]
    out = audio.pad_reflect_last_dim(x, 2, 2)
    assert out[0].tolist() == [2, 1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 7, 6]


def test_gqa_head_mapping():  # tensor_utils.rs:116-121
    x = np.arange(2 * 3, dtype=np.float32).reshape(1, 2, 3, 1)
    r = qwen3.repeat_kv(x, 2)
    assert [int(r[0, h, 0, 0]) // 3 for h in range(4)] == [0, 0, 1, 1]


def test_mrope_interleave_rule():  # rope.rs:454-476 for [24,20,20]
    f = np.stack([np.full((1, 1, 64), r, np.float32) for r in range(3)])
    out = rope.Qwen3VLTextRotaryEmbedding.apply_interleaved_mrope(f, [24, 20, 20])[0, 0]
    for j in range(64):
        want = 1 if (j % 3 == 1 and j < 60) else 2 if (j % 3 == 2 and j < 60) else 0
        assert out[j] == want


def test_causal_mask():  # tensor_utils.rs:89-95
    m = qwen3.prepare_causal_attention_mask(1, 4)[0, 0]
    for i in range(4):
        for j in range(4):
            assert (m[i, j] == -np.inf) == (j > i)


def test_inv_freq_formula():  # rope.rs:7-13
    f = rope.compute_default_rope_parameters(128, 1e6)
    assert f.shape == (64,) and f[0] == 1.0
    np.testing.assert_allclose(f[1], 1.0 / 1e6 ** (2 / 128), rtol=1e-6)


def test_activations_match_definitions():
    x = np.linspace(-4, 4, 33, dtype=np.float32)
    import math
    np.testing.assert_allclose(nn.gelu_erf(x), [0.5 * v * (1 + math.erf(v / math.sqrt(2))) for v in x], atol=1e-6)
    np.testing.assert_allclose(nn.silu(x), x / (1 + np.exp(-x)), atol=1e-6)


def test_generate_loop_semantics():  # generate.rs:115-159
    from oracle.generate import GenerationContext, generate_generic

    class Fake:
        def __init__(self, seq): self.seq, self.i, self.cleared = seq, 0, False
        def _logits(self):
            l = np.zeros(8, np.float32); l[self.seq[self.i]] = 1; self.i += 1; return l
        def forward_initial(self, ids, off, data): assert off == 0; return self._logits()
        def forward_step(self, ids, off): self.offs = getattr(self, "offs", []) + [off]; return self._logits()
        def clear_cache(self): self.cleared = True
        def stop_token_ids(self): return [7]

    m = Fake([7, 3, 7, 5])       # first token is EOS but never checked; the second EOS is pushed, then break
    ctx = GenerationContext(temperature=0.0, initial_seq_len=5, max_tokens=10)
    toks, _, _ = generate_generic(m, np.zeros((1, 5)), None, ctx)
    assert toks == [7, 3, 7] and m.cleared and m.offs == [5, 6]
