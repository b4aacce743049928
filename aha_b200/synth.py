"""Synthetic checkpoints: seeded random-init weights with the REAL tensor names and shapes the
reference's loaders read (`vb.pp(...)` calls at /root/reference/src/models/qwen3/model.rs:28-62,105-125,
qwen3vl/model.rs:54,61,129-132,201-202,291-313,389-410,848-860, qwen3_asr/model.rs:43-61,104-151,239-247,
317-326,375).  There is no network for real checkpoints; values are N(0, 0.02) (norm gains 1 + N(0, 0.02))
rounded to fp16 so that fp16 storage is lossless for both the CUDA path and the fp32 oracle.

Config dicts follow the HF `config.json` layout that the reference deserialises
(qwen3/config.rs, qwen3vl/config.rs, qwen3_asr/config.rs)."""
import copy
import numpy as np

# ----------------------------------------------------------------------------- configs
_TEXT_06 = dict(attention_bias=False, head_dim=128, hidden_act="silu", hidden_size=1024, intermediate_size=3072,
                num_attention_heads=16, num_hidden_layers=28, num_key_value_heads=8, rms_norm_eps=1e-6,
                rope_theta=1000000.0, tie_word_embeddings=True, vocab_size=151936, eos_token_id=151645)
_TEXT_TINY = dict(attention_bias=False, head_dim=128, hidden_act="silu", hidden_size=256, intermediate_size=512,
                  num_attention_heads=4, num_hidden_layers=2, num_key_value_heads=2, rms_norm_eps=1e-6,
                  rope_theta=1000000.0, tie_word_embeddings=True, vocab_size=1024, eos_token_id=1023)
_MROPE = dict(rope_type="default", mrope_section=[24, 20, 20], mrope_interleaved=True)

QWEN3 = {
    "tiny": dict(_TEXT_TINY),
    # Qwen3-VL-2B text-stack row shapes (H=2048, I=6144, 16/8 heads) with 2 layers and a small vocabulary: exercises the
    # production code paths of the decode kernels (4-row K=2048 stages, 1-row K=6144 stages) at test cost
    "mid": dict(_TEXT_06, hidden_size=2048, intermediate_size=6144, num_hidden_layers=2, vocab_size=4096, rope_theta=5000000.0,
                eos_token_id=4095),
    "tiny-untied": dict(_TEXT_TINY, tie_word_embeddings=False),
    # GQA groups other than 2 (every shipped Qwen3 size the fused kernel accepts has nh / nkv = 2): MHA and a group of 4
    "tiny-g1": dict(_TEXT_TINY, num_attention_heads=2, num_key_value_heads=2),
    "tiny-g4": dict(_TEXT_TINY, num_attention_heads=4, num_key_value_heads=1),
    "q0.6": dict(_TEXT_06),
}

_VIS_2B = dict(deepstack_visual_indexes=[5, 11, 17], depth=24, hidden_act="gelu_pytorch_tanh", hidden_size=1024,
               in_channels=3, intermediate_size=4096, num_heads=16, num_position_embeddings=2304,
               out_hidden_size=2048, patch_size=16, spatial_merge_size=2, temporal_patch_size=2)
_VIS_TINY = dict(deepstack_visual_indexes=[0, 2], depth=4, hidden_act="gelu_pytorch_tanh", hidden_size=128,
                 in_channels=3, intermediate_size=256, num_heads=2, num_position_embeddings=64,
                 out_hidden_size=256, patch_size=16, spatial_merge_size=2, temporal_patch_size=2)
# the Qwen3-VL-8B / 32B tower: hidden 1152 over 16 heads = head_dim 72, intermediate 4304 (neither tiles the tensor-core kernels: the
# library pads head slots to 128 and the intermediate to 4352 at upload, vision_model.cuh)
_VIS_8B = dict(_VIS_2B, depth=27, deepstack_visual_indexes=[8, 16, 24], hidden_size=1152, intermediate_size=4304, out_hidden_size=4096)
_VL_IDS = dict(image_token_id=151655, video_token_id=151656, vision_start_token_id=151652, vision_end_token_id=151653)

QWEN3VL = {
    "tiny": dict(image_token_id=1001, video_token_id=1002, vision_start_token_id=1003, vision_end_token_id=1004,
                 tie_word_embeddings=True,
                 text_config=dict(_TEXT_TINY, rope_theta=5000000.0, rope_scaling=dict(_MROPE)),
                 vision_config=dict(_VIS_TINY)),
    # head_dim 72 and an intermediate size that is not a multiple of 64, at test cost (K = 576 and the padded shapes tile the tcgen05 GEMM)
    "tiny-hd72": dict(image_token_id=1001, video_token_id=1002, vision_start_token_id=1003, vision_end_token_id=1004,
                      tie_word_embeddings=True,
                      text_config=dict(_TEXT_TINY, rope_theta=5000000.0, rope_scaling=dict(_MROPE)),
                      vision_config=dict(_VIS_TINY, depth=2, deepstack_visual_indexes=[0], hidden_size=576, num_heads=8,
                                         intermediate_size=1080)),
    "vl2": dict(_VL_IDS, tie_word_embeddings=True,
                text_config=dict(_TEXT_06, hidden_size=2048, intermediate_size=6144, rope_theta=5000000.0,
                                 rope_scaling=dict(_MROPE)),
                vision_config=dict(_VIS_2B)),
    # BASELINE.json config 5 says "7B"; the reference registry has 2B/4B/8B/32B only (model_mapping.rs:57-64) -> Qwen3-VL-8B: text stack
    # H 4096, 36 layers, 32 / 8 heads, I 12288; vision tower 27 blocks of hidden 1152 / 16 heads (head_dim 72), intermediate 4304.
    "vl8": dict(_VL_IDS, tie_word_embeddings=False,
                text_config=dict(_TEXT_06, hidden_size=4096, intermediate_size=12288, num_attention_heads=32,
                                 num_hidden_layers=36, rope_theta=5000000.0, rope_scaling=dict(_MROPE)),
                vision_config=dict(_VIS_8B)),
}

_AUD_06 = dict(activation_function="gelu", conv_chunksize=500, d_model=896, downsample_hidden_size=480,
               encoder_attention_heads=14, encoder_ffn_dim=3584, encoder_layers=18, n_window=50, n_window_infer=800,
               num_mel_bins=128, output_dim=1024, max_source_positions=1500)
_AUD_TINY = dict(activation_function="gelu", conv_chunksize=2, d_model=128, downsample_hidden_size=32,
                 encoder_attention_heads=2, encoder_ffn_dim=256, encoder_layers=2, n_window=50, n_window_infer=800,
                 num_mel_bins=128, output_dim=256, max_source_positions=1500)

QWEN3_ASR = {
    "tiny": dict(thinker_config=dict(audio_token_id=1001, audio_start_token_id=1002, audio_end_token_id=1003,
                                     audio_config=dict(_AUD_TINY),
                                     text_config=dict(_TEXT_TINY, rope_scaling=dict(_MROPE)))),
    "asr0.6": dict(thinker_config=dict(audio_token_id=151676, audio_start_token_id=151669, audio_end_token_id=151670,
                                       audio_config=dict(_AUD_06),
                                       text_config=dict(_TEXT_06, rope_scaling=dict(_MROPE)))),
}


def get_config(kind, preset):
    return copy.deepcopy({"qwen3": QWEN3, "qwen3vl": QWEN3VL, "qwen3_asr": QWEN3_ASR}[kind][preset])


# ----------------------------------------------------------------------------- weights
class _Gen:
    def __init__(self, seed, std=0.02, fast=False):
        self.rng = np.random.default_rng(seed)
        self.std = std
        self.fast = fast       # uniform with the same std instead of normal: 3x faster to draw (the 8-billion-parameter timing preset)
        self.w = {}

    def mat(self, name, *shape):
        n = int(np.prod(shape))
        if self.fast:
            a = self.rng.random(n, dtype=np.float32)
            a -= np.float32(0.5)
            a *= np.float32(self.std * 3.4641016)
        else:
            a = self.rng.standard_normal(n, dtype=np.float32)
            a *= np.float32(self.std)
        self.w[name] = a.astype(np.float16).reshape(shape)

    def gain(self, name, n):
        self.w[name] = (1.0 + self.std * self.rng.standard_normal(n, dtype=np.float32)).astype(np.float16)

    def bias(self, name, n):
        self.w[name] = (self.std * self.rng.standard_normal(n, dtype=np.float32)).astype(np.float16)

    def linear(self, prefix, out_f, in_f, bias):
        self.mat(prefix + ".weight", out_f, in_f)
        if bias:
            self.bias(prefix + ".bias", out_f)

    def lnorm(self, prefix, n):
        self.gain(prefix + ".weight", n)
        self.bias(prefix + ".bias", n)


def _text_layers(g, p, tc):
    H, I = tc["hidden_size"], tc["intermediate_size"]
    nh, nkv, hd = tc["num_attention_heads"], tc["num_key_value_heads"], tc["head_dim"]
    b = tc.get("attention_bias", False)
    g.mat(p + "embed_tokens.weight", tc["vocab_size"], H)
    for i in range(tc["num_hidden_layers"]):
        lp = f"{p}layers.{i}."
        g.linear(lp + "self_attn.q_proj", nh * hd, H, b)
        g.linear(lp + "self_attn.k_proj", nkv * hd, H, b)
        g.linear(lp + "self_attn.v_proj", nkv * hd, H, b)
        g.linear(lp + "self_attn.o_proj", H, nh * hd, b)
        g.gain(lp + "self_attn.q_norm.weight", hd)
        g.gain(lp + "self_attn.k_norm.weight", hd)
        g.linear(lp + "mlp.gate_proj", I, H, False)
        g.linear(lp + "mlp.up_proj", I, H, False)
        g.linear(lp + "mlp.down_proj", H, I, False)
        g.gain(lp + "input_layernorm.weight", H)
        g.gain(lp + "post_attention_layernorm.weight", H)
    g.gain(p + "norm.weight", H)


def make_qwen3(cfg, seed=0):
    g = _Gen(seed)
    _text_layers(g, "model.", cfg)
    if not cfg.get("tie_word_embeddings", False):
        g.mat("lm_head.weight", cfg["vocab_size"], cfg["hidden_size"])
    return g.w


def make_qwen3vl(cfg, seed=0):
    g = _Gen(seed, fast=cfg["text_config"]["hidden_size"] >= 4096)
    vc, tc = cfg["vision_config"], cfg["text_config"]
    Hv, Iv, p = vc["hidden_size"], vc["intermediate_size"], "model.visual."
    ps, tp, m = vc["patch_size"], vc["temporal_patch_size"], vc["spatial_merge_size"]
    g.mat(p + "patch_embed.proj.weight", Hv, vc["in_channels"], tp, ps, ps)
    g.bias(p + "patch_embed.proj.bias", Hv)
    g.mat(p + "pos_embed.weight", vc["num_position_embeddings"], Hv)
    for i in range(vc["depth"]):
        bp = f"{p}blocks.{i}."
        g.lnorm(bp + "norm1", Hv)
        g.lnorm(bp + "norm2", Hv)
        g.linear(bp + "attn.qkv", 3 * Hv, Hv, True)
        g.linear(bp + "attn.proj", Hv, Hv, True)
        g.linear(bp + "mlp.linear_fc1", Iv, Hv, True)
        g.linear(bp + "mlp.linear_fc2", Hv, Iv, True)

    def merger(mp, post):
        Hm = Hv * m * m
        g.lnorm(mp + "norm", Hm if post else Hv)
        g.linear(mp + "linear_fc1", Hm, Hm, True)
        g.linear(mp + "linear_fc2", vc["out_hidden_size"], Hm, True)

    merger(p + "merger.", False)
    for i in range(len(vc["deepstack_visual_indexes"])):
        merger(f"{p}deepstack_merger_list.{i}.", True)
    _text_layers(g, "model.language_model.", tc)
    if not cfg.get("tie_word_embeddings", False):
        g.mat("lm_head.weight", tc["vocab_size"], tc["hidden_size"])
    return g.w


def make_qwen3_asr(cfg, seed=0):
    g = _Gen(seed)
    tk = cfg["thinker_config"]
    ac, tc = tk["audio_config"], tk["text_config"]
    p, D, C = "thinker.audio_tower.", ac["d_model"], ac["downsample_hidden_size"]
    g.mat(p + "conv2d1.weight", C, 1, 3, 3); g.bias(p + "conv2d1.bias", C)
    g.mat(p + "conv2d2.weight", C, C, 3, 3); g.bias(p + "conv2d2.bias", C)
    g.mat(p + "conv2d3.weight", C, C, 3, 3); g.bias(p + "conv2d3.bias", C)
    fdim = ((((ac["num_mel_bins"] + 1) // 2 + 1) // 2 + 1) // 2)
    g.mat(p + "conv_out.weight", D, C * fdim)
    for i in range(ac["encoder_layers"]):
        lp = f"{p}layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            g.linear(lp + "self_attn." + n, D, D, True)
        g.lnorm(lp + "self_attn_layer_norm", D)
        g.linear(lp + "fc1", ac["encoder_ffn_dim"], D, True)
        g.linear(lp + "fc2", D, ac["encoder_ffn_dim"], True)
        g.lnorm(lp + "final_layer_norm", D)
    g.lnorm(p + "ln_post", D)
    g.linear(p + "proj1", D, D, True)
    g.linear(p + "proj2", ac["output_dim"], D, True)
    _text_layers(g, "thinker.model.", tc)
    if not tc.get("tie_word_embeddings", False):
        g.mat("thinker.lm_head.weight", tc["vocab_size"], tc["hidden_size"])
    return g.w


def make_weights(kind, cfg, seed=0):
    return {"qwen3": make_qwen3, "qwen3vl": make_qwen3vl, "qwen3_asr": make_qwen3_asr}[kind](cfg, seed)


# ----------------------------------------------------------------------------- synthetic inputs
# BASELINE.json's full-size workloads (configs 3, 2, 4); tests/golden/make_golden_full.py, tests/test_fullsize_gpu.py and
# bench.py all build their inputs from these
FULL_VL2_IMAGE = (1088, 1920)    # img_smart_resize(1080, 1920, 32, ...) -- SURVEY 8c known answer
FULL_VL2_TEXT = 512
FULL_Q06_PROMPT = 1920
FULL_ASR_SECONDS = 30.0


def synth_text_ids(n, vocab, seed, avoid=()):
    """Uniform ids in [0, vocab) avoiding the given special ids."""
    rng = np.random.default_rng(seed)
    ids = rng.integers(0, vocab, size=n, dtype=np.int64)
    avoid = set(int(a) for a in avoid)
    for i in range(n):
        while int(ids[i]) in avoid:
            ids[i] = (ids[i] + 7) % vocab
    return ids.astype(np.uint32)


def synth_image(h, w, seed=1):
    """Uniform-noise uint8 HWC image (already a multiple of 32 so the reference's resize is the identity)."""
    return np.random.default_rng(seed).integers(0, 256, size=(h, w, 3), dtype=np.uint8)


def synth_audio(seconds, sr=16000, seed=2):
    """Sum of 5 sines + N(0, 0.01) noise, peak 0.9."""
    rng = np.random.default_rng(seed)
    n = int(seconds * sr)
    t = np.arange(n, dtype=np.float64) / sr
    x = sum(np.sin(2 * np.pi * f * t + ph) for f, ph in zip((220.0, 440.0, 880.0, 1760.0, 3520.0), rng.uniform(0, 6.28, 5)))
    x = x + 0.01 * rng.standard_normal(n)
    x = 0.9 * x / np.max(np.abs(x))
    return x.astype(np.float32)


def vl_prompt_ids(cfg, grid_thw, n_text, seed=3):
    """<|vision_start|> + (t*h*w/merge^2) x <|image_pad|> + <|vision_end|> per image, then n_text synthetic ids."""
    m2 = cfg["vision_config"]["spatial_merge_size"] ** 2
    special = (cfg["image_token_id"], cfg["video_token_id"], cfg["vision_start_token_id"], cfg["vision_end_token_id"])
    ids = []
    for t, h, w in np.asarray(grid_thw).tolist():
        ids += [cfg["vision_start_token_id"]] + [cfg["image_token_id"]] * (t * h * w // m2) + [cfg["vision_end_token_id"]]
    V = cfg["text_config"]["vocab_size"]
    ids += synth_text_ids(n_text, min(V, 151000), seed, avoid=special).tolist()
    return np.asarray(ids, dtype=np.uint32)


def vl_video_prompt_ids(cfg, video_grid_thw, n_stamp=3, seed=5):
    """One video the way Qwen3VLProcessor lays it out (qwen3vl/processor.rs:447-571): per temporal grid step a few timestamp text ids,
    then <|vision_start|> + (h*w/merge^2) x <|video_pad|> + <|vision_end|>."""
    m2 = cfg["vision_config"]["spatial_merge_size"] ** 2
    special = (cfg["image_token_id"], cfg["video_token_id"], cfg["vision_start_token_id"], cfg["vision_end_token_id"])
    V = cfg["text_config"]["vocab_size"]
    ids = []
    k = 0
    for t, h, w in np.asarray(video_grid_thw).tolist():
        for _ in range(t):
            ids += synth_text_ids(n_stamp, min(V, 151000), seed + k, avoid=special).tolist()
            ids += [cfg["vision_start_token_id"]] + [cfg["video_token_id"]] * (h * w // m2) + [cfg["vision_end_token_id"]]
            k += 1
    return np.asarray(ids, dtype=np.uint32)


def asr_audio_tokens(n_frames):
    """Audio tokens of `n_frames` log-mel frames: 13 per full 100-frame chunk + three stride-2 convolutions over the rest
    (get_feat_extract_output_lengths, /root/reference/src/models/qwen3_asr/processor.rs:187-195)."""
    leave = n_frames % 100
    tail = ((((leave - 1) // 2 + 1) - 1) // 2 + 1 - 1) // 2 + 1 if leave > 0 else 0
    return tail + (n_frames // 100) * 13


def asr_prompt_ids(cfg, n_audio_tokens, n_text=8, seed=4):
    tk = cfg["thinker_config"]
    special = (tk["audio_token_id"], tk["audio_start_token_id"], tk["audio_end_token_id"])
    V = tk["text_config"]["vocab_size"]
    pre = synth_text_ids(n_text // 2, min(V, 151000), seed, avoid=special).tolist()
    post = synth_text_ids(n_text - n_text // 2, min(V, 151000), seed + 1, avoid=special).tolist()
    ids = pre + [tk["audio_start_token_id"]] + [tk["audio_token_id"]] * n_audio_tokens + [tk["audio_end_token_id"]] + post
    return np.asarray(ids, dtype=np.uint32)
