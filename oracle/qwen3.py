"""Qwen3 dense decoder -- restates /root/reference/src/models/qwen3/model.rs and the shared
layers in src/models/common/modules.rs, src/utils/tensor_utils.rs.

Weights are a dict name -> np.ndarray using the checkpoint's tensor names."""
import numpy as np

from . import nn
from .rope import RoPE, apply_rotary_pos_emb

F32 = np.float32


def prepare_causal_attention_mask(b_size, tgt_len, seqlen_offset=0):
    """tensor_utils.rs:78-106: (b,1,T,T+off) f32, -inf where col > row."""
    ar = np.arange(tgt_len)
    m = np.where(ar[None, :] > ar[:, None], -np.inf, 0.0).astype(F32)
    if seqlen_offset > 0:
        m = np.concatenate([np.zeros((tgt_len, seqlen_offset), F32), m], axis=-1)
    return np.broadcast_to(m[None, None], (b_size, 1, tgt_len, tgt_len + seqlen_offset))


def repeat_kv(xs, n_rep):
    """tensor_utils.rs:108-124: cat n_rep copies on dim 2 then reshape => head h <-> kv h // n_rep."""
    if n_rep == 1:
        return xs
    b, nkv, s, hd = xs.shape
    return np.concatenate([xs] * n_rep, axis=2).reshape(b, nkv * n_rep, s, hd)


def eager_attention_forward(q, k, v, num_key_value_groups, attention_mask, scaling):
    """modules.rs:757-813 (non-flash branch).  q (b,H,S,hd); k,v (b,Hkv,T,hd) -> (b,S,H,hd)."""
    if num_key_value_groups is not None:
        k = repeat_kv(k, num_key_value_groups)
        v = repeat_kv(v, num_key_value_groups)
    q = np.ascontiguousarray(q)
    k = np.ascontiguousarray(k)
    v = np.ascontiguousarray(v)
    w = np.matmul(q, np.ascontiguousarray(np.swapaxes(k, -2, -1)))
    w = (w * F32(scaling)).astype(F32)
    if attention_mask is not None:
        w = w + attention_mask.astype(w.dtype)
    w = nn.softmax_last_dim(w)
    o = np.matmul(w, v)
    return np.ascontiguousarray(np.swapaxes(o, 1, 2))


class GateUpDownMLP:
    """modules.rs:48-87: down(act(gate(x)) * up(x)), no bias."""

    def __init__(self, w, prefix, act):
        self.gate = w[prefix + "gate_proj.weight"]
        self.up = w[prefix + "up_proj.weight"]
        self.down = w[prefix + "down_proj.weight"]
        self.act = nn.activation(act)

    def forward(self, x):
        return nn.linear(self.act(nn.linear(x, self.gate)) * nn.linear(x, self.up), self.down)


class QKNormAttention:
    """modules.rs:447-584."""

    def __init__(self, w, prefix, num_heads, head_dim, num_kv_heads, eps, bias=False):
        g = lambda n: w[prefix + n]
        gb = (lambda n: w.get(prefix + n)) if bias else (lambda n: None)
        self.q, self.k, self.v, self.o = g("q_proj.weight"), g("k_proj.weight"), g("v_proj.weight"), g("o_proj.weight")
        self.qb, self.kb, self.vb, self.ob = gb("q_proj.bias"), gb("k_proj.bias"), gb("v_proj.bias"), gb("o_proj.bias")
        self.q_norm, self.k_norm = g("q_norm.weight"), g("k_norm.weight")
        self.nh, self.nkv, self.hd, self.eps = num_heads, num_kv_heads, head_dim, eps
        self.groups = num_heads // num_kv_heads
        self.scaling = 1.0 / np.sqrt(np.float64(head_dim))
        self.kv_cache = None

    def forward(self, xs, cos, sin, mask):
        b, s, _ = xs.shape
        q = nn.linear(xs, self.q, self.qb).reshape(b, s, self.nh, self.hd)
        q = np.swapaxes(nn.rms_norm(q, self.q_norm, self.eps), 1, 2)
        k = nn.linear(xs, self.k, self.kb).reshape(b, s, self.nkv, self.hd)
        k = np.swapaxes(nn.rms_norm(k, self.k_norm, self.eps), 1, 2)
        v = np.swapaxes(nn.linear(xs, self.v, self.vb).reshape(b, s, self.nkv, self.hd), 1, 2)
        q, k = apply_rotary_pos_emb(q, k, cos, sin)
        if self.kv_cache is not None:  # modules.rs:558-566 -- whole-cache cat every step
            pk, pv = self.kv_cache
            k = np.concatenate([pk, k], axis=2)
            v = np.concatenate([pv, v], axis=2)
        self.kv_cache = (k, v)
        o = eager_attention_forward(q, k, v, self.groups, mask, self.scaling)
        return nn.linear(o.reshape(b, s, self.nh * self.hd), self.o, self.ob)

    def clear_kv_cache(self):
        self.kv_cache = None


class Qwen3DecoderLayer:
    """qwen3/model.rs:19-91."""

    def __init__(self, cfg, w, prefix):
        self.attn = QKNormAttention(w, prefix + "self_attn.", cfg["num_attention_heads"], cfg["head_dim"],
                                    cfg["num_key_value_heads"], cfg["rms_norm_eps"],
                                    cfg.get("attention_bias", False))
        self.mlp = GateUpDownMLP(w, prefix + "mlp.", cfg.get("hidden_act", "silu"))
        self.ln1 = w[prefix + "input_layernorm.weight"]
        self.ln2 = w[prefix + "post_attention_layernorm.weight"]
        self.eps = cfg["rms_norm_eps"]

    def forward(self, xs, cos, sin, mask):
        xs = xs + self.attn.forward(nn.rms_norm(xs, self.ln1, self.eps), cos, sin, mask)
        xs = xs + self.mlp.forward(nn.rms_norm(xs, self.ln2, self.eps))
        return xs.astype(F32)

    def clear_kv_cache(self):
        self.attn.clear_kv_cache()


class Qwen3Model:
    """qwen3/model.rs:94-214 incl. `impl InferenceModel`."""

    def __init__(self, cfg, w, eos_ids=()):
        self.cfg = cfg
        p = "model." if "model.embed_tokens.weight" in w else ""  # model.rs:105-109
        self.embed = w[p + "embed_tokens.weight"]
        self.layers = [Qwen3DecoderLayer(cfg, w, f"{p}layers.{i}.") for i in range(cfg["num_hidden_layers"])]
        self.norm = w[p + "norm.weight"]
        self.rotary = RoPE(cfg["head_dim"], cfg["rope_theta"])
        self.lm_head = self.embed if cfg.get("tie_word_embeddings", False) else w["lm_head.weight"]
        self._stop = list(eos_ids)
        self.trace = None  # optional list collecting per-layer hidden states (tests)

    def forward_hidden(self, input_ids=None, inputs_embeds=None, seqlen_offset=0):
        if input_ids is None and inputs_embeds is None:
            raise ValueError("You must specify exactly one of input_ids or inputs_embeds")
        x = inputs_embeds if inputs_embeds is not None else nn.embedding(input_ids, self.embed)
        b, s, _ = x.shape
        mask = prepare_causal_attention_mask(b, s, 0) if s > 1 else None  # model.rs:164-175
        cos, sin = self.rotary.forward(seqlen_offset, s)
        for layer in self.layers:
            x = layer.forward(x, cos, sin, mask)
            if self.trace is not None:
                self.trace.append(x.copy())
        x = nn.rms_norm(x, self.norm, self.cfg["rms_norm_eps"])
        return x[:, s - 1:s, :]

    def forward(self, input_ids=None, inputs_embeds=None, seqlen_offset=0):
        return nn.linear(self.forward_hidden(input_ids, inputs_embeds, seqlen_offset), self.lm_head)

    # InferenceModel (common/mod.rs:25-45)
    def forward_initial(self, input_ids, seqlen_offset, data=None):
        return self.forward_step(input_ids, seqlen_offset)

    def forward_step(self, input_ids, seqlen_offset):
        return self.forward(np.asarray(input_ids).reshape(1, -1), None, seqlen_offset)

    def clear_cache(self):
        for l in self.layers:
            l.clear_kv_cache()

    def stop_token_ids(self):
        return list(self._stop)
