"""Qwen3-ASR -- restates /root/reference/src/models/qwen3_asr/model.rs and
src/models/common/modules.rs:127-242 (NaiveAttention)."""
import numpy as np

from . import nn
from .audio import get_feat_extract_output_lengths
from .qwen3 import Qwen3DecoderLayer, eager_attention_forward, prepare_causal_attention_mask
from .rope import Qwen3VLTextRotaryEmbedding, SinusoidalPositionEncoderCat

F32 = np.float32


class Qwen3ASRAudioEncoderLayer:
    """model.rs:32-83: pre-LN MHA (biases, no RoPE, no mask) + pre-LN FFN."""

    def __init__(self, ac, w, prefix):
        g = lambda n: w[prefix + n]
        self.nh = ac["encoder_attention_heads"]
        self.hd = ac["d_model"] // self.nh
        self.q, self.qb = g("self_attn.q_proj.weight"), g("self_attn.q_proj.bias")
        self.k, self.kb = g("self_attn.k_proj.weight"), g("self_attn.k_proj.bias")
        self.v, self.vb = g("self_attn.v_proj.weight"), g("self_attn.v_proj.bias")
        self.o, self.ob = g("self_attn.out_proj.weight"), g("self_attn.out_proj.bias")
        self.ln1w, self.ln1b = g("self_attn_layer_norm.weight"), g("self_attn_layer_norm.bias")
        self.ln2w, self.ln2b = g("final_layer_norm.weight"), g("final_layer_norm.bias")
        self.fc1, self.fc1b, self.fc2, self.fc2b = g("fc1.weight"), g("fc1.bias"), g("fc2.weight"), g("fc2.bias")
        self.act = nn.activation(ac.get("activation_function", "gelu"))

    def forward(self, xs):
        b, s, _ = xs.shape
        h = nn.layer_norm(xs, self.ln1w, self.ln1b, 1e-5)
        q = np.swapaxes(nn.linear(h, self.q, self.qb).reshape(b, s, self.nh, self.hd), 1, 2)
        k = np.swapaxes(nn.linear(h, self.k, self.kb).reshape(b, s, self.nh, self.hd), 1, 2)
        v = np.swapaxes(nn.linear(h, self.v, self.vb).reshape(b, s, self.nh, self.hd), 1, 2)
        o = eager_attention_forward(q, k, v, 1, None, 1.0 / np.sqrt(np.float64(self.hd)))
        res = nn.linear(o.reshape(b, s, -1), self.o, self.ob) + xs
        h = nn.layer_norm(res, self.ln2w, self.ln2b, 1e-5)
        return (nn.linear(self.act(nn.linear(h, self.fc1, self.fc1b)), self.fc2, self.fc2b) + res).astype(F32)


class Qwen3ASRAudioEncoder:
    """model.rs:85-227."""

    def __init__(self, ac, w, prefix="thinker.audio_tower."):
        self.ac = ac
        self.n_window = ac["n_window"]
        self.pe = SinusoidalPositionEncoderCat(ac["d_model"])
        self.layers = [Qwen3ASRAudioEncoderLayer(ac, w, f"{prefix}layers.{i}.") for i in range(ac["encoder_layers"])]
        g = lambda n: w[prefix + n]
        self.convs = [(g(f"conv2d{i}.weight"), g(f"conv2d{i}.bias")) for i in (1, 2, 3)]
        self.conv_out = g("conv_out.weight")
        self.lnw, self.lnb = g("ln_post.weight"), g("ln_post.bias")
        self.p1, self.p1b, self.p2, self.p2b = g("proj1.weight"), g("proj1.bias"), g("proj2.weight"), g("proj2.bias")
        self.act = nn.activation(ac.get("activation_function", "gelu"))
        self.conv_chunksize = ac.get("conv_chunksize", 500)
        self.conv_act = nn.gelu_tanh  # Tensor::gelu() == tanh approx (model.rs:200-202); HF uses erf here (tests swap it to cross-check the rest)
        self.trace = None

    def forward(self, xs):
        """xs: (n_mels, T) -> (n_tokens, output_dim)."""
        T = xs.shape[1]
        cw = self.n_window * 2
        lens = [cw] * (T // cw)
        if T % cw:
            lens.append(T % cw)
        xt = xs.T
        chunks, o = [], 0
        for L in lens:
            c = xt[o:o + L]
            o += L
            if L < cw:
                c = np.concatenate([c, np.zeros((cw - L, c.shape[1]), F32)], axis=0)
            chunks.append(c)
        feat = np.swapaxes(np.stack(chunks, 0), 1, 2)[:, None]  # (B,1,mel,cw)
        lens_after = [get_feat_extract_output_lengths(L) for L in lens]
        total = sum(lens_after)
        outs = []
        for s in range(0, feat.shape[0], self.conv_chunksize):
            e = feat[s:s + self.conv_chunksize]
            for cwt, cb in self.convs:
                e = self.conv_act(nn.conv2d(e, cwt, cb, 2, 1))
            outs.append(e)
        e = np.concatenate(outs, 0)
        b, c, f, t = e.shape
        e = np.ascontiguousarray(np.transpose(e, (0, 3, 1, 2))).reshape(b, t, c * f)
        e = nn.linear(e, self.conv_out)
        e = self.pe.forward(e, 0).astype(F32)  # positions restart at 0 per chunk
        h = e.reshape(b * t, -1)[:total][None]
        if self.trace is not None:
            self.trace.append(("conv", h[0].copy()))
        for i, layer in enumerate(self.layers):
            h = layer.forward(h)  # whole sequence, mask=None (model.rs:218-220)
            if self.trace is not None:
                self.trace.append((f"layer{i}", h[0].copy()))
        h = nn.layer_norm(h[0], self.lnw, self.lnb, 1e-5)
        return nn.linear(self.act(nn.linear(h, self.p1, self.p1b)), self.p2, self.p2b)


class Qwen3ASRModel:
    """model.rs:229-425 (thinker text model + thinker + InferenceModel)."""

    def __init__(self, cfg, w, eos_ids=()):
        tk = cfg["thinker_config"]
        self.tc = tk["text_config"]
        self.audio = Qwen3ASRAudioEncoder(tk["audio_config"], w)
        p = "thinker.model."
        self.embed = w[p + "embed_tokens.weight"]
        self.layers = [Qwen3DecoderLayer(self.tc, w, f"{p}layers.{i}.") for i in range(self.tc["num_hidden_layers"])]
        self.norm = w[p + "norm.weight"]
        self.rotary = Qwen3VLTextRotaryEmbedding(self.tc["head_dim"], self.tc["rope_theta"])
        self.mrope_section = list(self.tc["rope_scaling"]["mrope_section"])
        self.audio_token_id = tk["audio_token_id"]
        self.lm_head = self.embed if self.tc.get("tie_word_embeddings", False) else w["thinker.lm_head.weight"]
        self._stop = list(eos_ids)

    def forward(self, input_ids, seqlen_offset, input_features=None):
        ids = np.asarray(input_ids).reshape(1, -1)
        x = nn.embedding(ids, self.embed)
        if input_features is not None:
            feat = self.audio.forward(input_features)
            mask = ids[0] == self.audio_token_id
            if int(mask.sum()) != feat.shape[0]:  # model.rs:348-354
                raise ValueError(f"n_audio_tokens num: {int(mask.sum())} not equal to audio_feature len: {feat.shape[0]}")
            x = x.copy()
            x[0, np.nonzero(mask)[0]] = feat
        b, s, _ = x.shape
        pos = np.broadcast_to(np.arange(seqlen_offset, seqlen_offset + s, dtype=np.int64)[None, None], (3, b, s))
        cos, sin = self.rotary.forward_asr(pos, self.mrope_section)
        mask = prepare_causal_attention_mask(b, s, 0) if s > 1 else None
        for layer in self.layers:
            x = layer.forward(x, cos, sin, mask)
        x = nn.rms_norm(x, self.norm, self.tc["rms_norm_eps"])
        return nn.linear(x[:, s - 1:s, :], self.lm_head)

    def forward_initial(self, input_ids, seqlen_offset, data):
        if data is None or len(data) != 1:
            raise ValueError("Qwen3 asr process data error")
        return self.forward(input_ids, seqlen_offset, data[0])

    def forward_step(self, input_ids, seqlen_offset):
        return self.forward(input_ids, seqlen_offset, None)

    def clear_cache(self):
        for l in self.layers:
            l.clear_kv_cache()

    def stop_token_ids(self):
        return list(self._stop)
