"""Qwen3-VL -- restates /root/reference/src/models/qwen3vl/{model,processor}.rs,
src/utils/img_utils.rs:272-331, src/utils/mod.rs:392-405, src/models/common/gguf.rs:384-390."""
import numpy as np

from . import nn
from .qwen3 import Qwen3DecoderLayer, eager_attention_forward, prepare_causal_attention_mask
from .rope import Qwen2_5VisionRotaryEmbedding, Qwen3VLTextRotaryEmbedding, apply_rotary_pos_emb_vision

F32 = np.float32


# ----------------------------------------------------------------------------- host preprocessing
def round_by_factor(num, factor):
    """utils/mod.rs:392-395 (f32 round = half away from zero)."""
    q = np.float32(num) / np.float32(factor)
    return int(np.floor(q + np.float32(0.5))) * factor if q >= 0 else int(np.ceil(q - np.float32(0.5))) * factor


def floor_by_factor(num, factor):
    return int(np.floor(np.float32(num) / np.float32(factor))) * factor


def ceil_by_factor(num, factor):
    return int(np.ceil(np.float32(num) / np.float32(factor))) * factor


def img_smart_resize(img_h, img_w, factor, min_pixels, max_pixels):
    """img_utils.rs:297-331; returns (height, width)."""
    if max(img_h, img_w) // min(img_h, img_w) > 200:
        raise ValueError("absolute aspect ratio mush be smaller than 200")
    h_bar = max(factor, round_by_factor(img_h, factor))
    w_bar = max(factor, round_by_factor(img_w, factor))
    if h_bar * w_bar > max_pixels:
        beta = np.sqrt(np.float32(img_h * img_w) / np.float32(max_pixels), dtype=F32)
        h_bar = max(factor, floor_by_factor(np.float32(img_h) / beta, factor))
        w_bar = max(factor, floor_by_factor(np.float32(img_w) / beta, factor))
    elif h_bar * w_bar < min_pixels:
        beta = np.sqrt(np.float32(min_pixels) / np.float32(img_h * img_w), dtype=F32)
        h_bar = ceil_by_factor(np.float32(img_h) * beta, factor)
        w_bar = ceil_by_factor(np.float32(img_w) * beta, factor)
    return h_bar, w_bar


def img_transform(img_u8_hwc, mean, std):
    """img_utils.rs:272-294: u8 HWC -> f32 CHW, *(1/255), (x-mean)/std."""
    x = np.transpose(img_u8_hwc, (2, 0, 1)).astype(F32) * F32(1.0 / 255.0)
    return ((x - np.asarray(mean, F32).reshape(3, 1, 1)) / np.asarray(std, F32).reshape(3, 1, 1)).astype(F32)


def _catmullrom_kernel(x):
    """`image` crate 0.25 imageops::sample::catmullrom_kernel = bc_cubic_spline(x, 0.0, 0.5): all f32, no FMA."""
    a = np.abs(np.asarray(x, F32))
    a2 = (a * a).astype(F32)
    a3 = (a2 * a).astype(F32)
    near = ((F32(9.0) * a3).astype(F32) + (F32(-15.0) * a2).astype(F32)).astype(F32) + F32(6.0)
    far = (((F32(-3.0) * a3).astype(F32) + (F32(15.0) * a2).astype(F32)).astype(F32) + (F32(-24.0) * a).astype(F32)).astype(F32) + F32(12.0)
    k = np.where(a < 1, near, np.where(a < 2, far, F32(0))).astype(F32)
    return (k / F32(6.0)).astype(F32)


def _resize_taps(out_i, in_n, out_n, support=2.0):
    """one output coordinate of imageops::{vertical,horizontal}_sample: (left, normalised weights)."""
    ratio = F32(in_n) / F32(out_n)
    sratio = F32(1.0) if ratio < 1 else ratio
    src_support = F32(support) * sratio
    inp = (F32(out_i) + F32(0.5)) * ratio
    left = int(min(max(int(np.floor(inp - src_support)), 0), in_n - 1))
    right = int(min(max(int(np.ceil(inp + src_support)), left + 1), in_n))
    inp = F32(inp - F32(0.5))
    ws = _catmullrom_kernel(((np.arange(left, right, dtype=F32) - inp) / sratio).astype(F32))
    s = F32(0)
    for w in ws:
        s = F32(s + w)
    return left, (ws / s).astype(F32)


def resize_exact_catmullrom(img_u8_hwc, new_h, new_w):
    """DynamicImage::resize_exact(new_w, new_h, FilterType::CatmullRom) for an RGB8 image (qwen3vl/processor.rs:167).
    THIRD-PARTY arithmetic (`image` 0.25.10, not under /root/reference), restated from imageops::resize: a copy when the size
    already matches, else vertical_sample into f32 followed by horizontal_sample back to u8 (clamp to [0, 255], round half away
    from zero); taps accumulate left to right in f32."""
    img = np.asarray(img_u8_hwc, np.uint8)
    h, w, _ = img.shape
    if (new_h, new_w) == (h, w):
        return img.copy()
    tmp = np.zeros((new_h, w, 3), F32)
    for oy in range(new_h):
        left, ws = _resize_taps(oy, h, new_h)
        t = np.zeros((w, 3), F32)
        for i, wt in enumerate(ws):
            t = (t + (img[left + i].astype(F32) * wt).astype(F32)).astype(F32)
        tmp[oy] = t
    out = np.zeros((new_h, new_w, 3), np.uint8)
    for ox in range(new_w):
        left, ws = _resize_taps(ox, w, new_w)
        t = np.zeros((new_h, 3), F32)
        for i, wt in enumerate(ws):
            t = (t + (tmp[:, left + i].astype(F32) * wt).astype(F32)).astype(F32)
        t = np.clip(t, F32(0), F32(255))
        out[:, ox] = np.floor(t + F32(0.5)).astype(np.uint8)      # t >= 0: round half away from zero == floor(t + 0.5)
    return out


def expand_placeholders(ids, token_id, counts):
    """processor.rs:386-399 / qwen3_asr/processor.rs:93-97 on token ids: the i-th occurrence of the pad token becomes counts[i] copies."""
    out, k = [], 0
    for t in np.asarray(ids).reshape(-1).tolist():
        if t == token_id and k < len(counts):
            out += [t] * int(counts[k]); k += 1
        else:
            out.append(t)
    return np.asarray(out, np.uint32)


def process_vision_tensor(img_tchw, patch_size=16, temporal_patch_size=2, merge_size=2):
    """processor.rs:174-227: (t,c,h,w) -> (grid_t*grid_h*grid_w, c*tp*p*p), grid_thw (1,3)."""
    t = img_tchw.shape[0]
    if t % temporal_patch_size != 0:
        rep = temporal_patch_size - t % temporal_patch_size
        img_tchw = np.concatenate([img_tchw, np.repeat(img_tchw[t - 1:t], rep, axis=0)], axis=0)
    c = img_tchw.shape[1]
    gt = img_tchw.shape[0] // temporal_patch_size
    gh = img_tchw.shape[2] // patch_size
    gw = img_tchw.shape[3] // patch_size
    x = img_tchw.reshape(gt, temporal_patch_size, c, gh // merge_size, merge_size, patch_size,
                         gw // merge_size, merge_size, patch_size)
    x = np.transpose(x, (0, 3, 6, 4, 7, 2, 1, 5, 8))
    x = np.ascontiguousarray(x).reshape(gt * gh * gw, c * temporal_patch_size * patch_size * patch_size)
    return x, np.array([[gt, gh, gw]], dtype=np.uint32)


def process_image(img_u8_hwc, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), patch_size=16, temporal_patch_size=2,
                  merge_size=2, min_pixels=65536, max_pixels=16777216):
    """processor.rs:151-172,229-251 for one image: img_smart_resize -> CatmullRom resize_exact -> img_transform; the frame
    is duplicated (T=2) -- processor.rs:240."""
    h, w = img_u8_hwc.shape[:2]
    rh, rw = img_smart_resize(h, w, patch_size * merge_size, min_pixels, max_pixels)
    if (rh, rw) != (h, w):
        img_u8_hwc = resize_exact_catmullrom(img_u8_hwc, rh, rw)
    x = img_transform(img_u8_hwc, mean, std)[None]
    x = np.concatenate([x, x], axis=0)
    return process_vision_tensor(x, patch_size, temporal_patch_size, merge_size)


# ----------------------------------------------------------------------------- video half of the processor
def video_smart_resize(num_frames, height, width, temporal_factor, factor, min_pixels, max_pixels, video_ratio=None):
    """utils/video_utils.rs:9-59; returns (height, width).  The pixel products are u32 in the reference (a release build wraps)."""
    u32 = lambda v: int(v) & 0xFFFFFFFF
    if num_frames < temporal_factor:
        raise ValueError(f"{num_frames} must be larger than temporal_factor {temporal_factor}")
    if height < factor or width < factor:
        raise ValueError(f"height:{height} or width:{width} must be larger than factor:{factor}")
    if max(height, width) // min(height, width) > 200:
        raise ValueError("absolute aspect ratio mush be smaller than 200, got %d" % (max(height, width) // min(height, width)))
    image_factor = factor
    if video_ratio is not None:
        image_factor = int(np.lcm(image_factor, video_ratio))
    h_bar = round_by_factor(height, image_factor)
    w_bar = round_by_factor(width, image_factor)
    t_bar = round_by_factor(num_frames, temporal_factor)
    if u32(t_bar * h_bar * w_bar) > max_pixels:
        beta = np.sqrt(F32(u32(num_frames * height * width)) / F32(max_pixels), dtype=F32)
        h_bar = max(image_factor, floor_by_factor(F32(height) / beta, image_factor))
        w_bar = max(image_factor, floor_by_factor(F32(width) / beta, image_factor))
    elif u32(t_bar * h_bar * w_bar) < min_pixels:
        beta = np.sqrt(F32(min_pixels) / F32(u32(num_frames * height * width)), dtype=F32)
        h_bar = ceil_by_factor(F32(height) * beta, image_factor)
        w_bar = ceil_by_factor(F32(width) * beta, image_factor)
    return h_bar, w_bar


def _round_f32(x):
    """f32::round (half away from zero) of a non-negative f32."""
    return int(np.floor(float(F32(x)) + 0.5))   # exact in f64 for any f32 of this magnitude


def video_sample_frames(total_frames, rate_num, rate_den, fps, min_frames, max_frames):
    """get_video_data's sampling (processor.rs:481-491, 526-527): returns (nframes, kept frame indices)."""
    rate = F32(rate_num) / F32(rate_den)
    nframes = _round_f32(F32(total_frames) / rate * F32(fps))
    nframes = min(min(max(nframes, min_frames), max_frames), total_frames)
    interval = _round_f32(F32(total_frames) / F32(nframes))
    return nframes, [f for f in range(total_frames) if f % interval == 0]


def calculate_timestamps(frames_indices, fps, t_merge_size):
    """processor.rs:282-307."""
    idx = list(frames_indices)
    if len(idx) % t_merge_size != 0:
        idx += [idx[-1]] * (t_merge_size - len(idx) % t_merge_size)
    ts = [F32(x) / F32(fps) for x in idx]
    return [F32((ts[i] + ts[i + t_merge_size - 1]) / F32(2.0)) for i in range(0, len(ts), t_merge_size)]


def format_timestamp(t):
    """format!("<{:.1} seconds>", t): the exactly rounded decimal of the f32 value."""
    return "<%.1f seconds>" % float(F32(t))


def process_video(frames_u8_thwc, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), patch_size=16, temporal_patch_size=2, merge_size=2):
    """process_videos for one clip (processor.rs:253-280): the (t, c, h, w) tensor get_video_data stacks from RGB24 frames,
    * 1/255, (x - mean) / std, then process_vision_tensor."""
    x = np.transpose(np.asarray(frames_u8_thwc), (0, 3, 1, 2)).astype(F32) * F32(1.0 / 255.0)
    x = (x - np.asarray(mean, F32).reshape(1, 3, 1, 1)) / np.asarray(std, F32).reshape(1, 3, 1, 1)
    return process_vision_tensor(x.astype(F32), patch_size, temporal_patch_size, merge_size)


def expand_video_placeholders_text(text, video_grid_thw, stamps_per_video, merge_size=2, video_token="<|video_pad|>",
                                   vision_start="<|vision_start|>", vision_end="<|vision_end|>"):
    """processor.rs:404-437 verbatim, on the STRING the reference edits (stamps_per_video[i] = calculate_timestamps of video i)."""
    merge_length = merge_size ** 2
    index = 0
    while video_token in text:
        t, h, w = [int(v) for v in video_grid_thw[index]]
        frame_seqlen = h * w // merge_length
        ph = ""
        for frame_idx in range(t):
            ph += format_timestamp(stamps_per_video[index][frame_idx])
            ph += vision_start + "<|placeholder|>" * frame_seqlen + vision_end
        three = vision_start + video_token + vision_end
        text = text.replace(three, ph, 1) if three in text else text.replace(video_token, ph, 1)
        index += 1
    return text.replace("<|placeholder|>", video_token)


def linspace(start, end, steps):
    """tensor_utils.rs:354-365 (f32: start + i*step)."""
    if steps == 1:
        return np.array([start], dtype=F32)
    step = (F32(end) - F32(start)) / F32(steps - 1)
    return (F32(start) + np.arange(steps, dtype=F32) * step).astype(F32)


# ----------------------------------------------------------------------------- vision tower
class Qwen3VLVisionPatchMerger:
    """model.rs:106-185."""

    def __init__(self, vc, w, prefix, use_postshuffle_norm):
        self.hidden = vc["hidden_size"] * vc["spatial_merge_size"] ** 2
        self.post = use_postshuffle_norm
        self.nw, self.nb = w[prefix + "norm.weight"], w[prefix + "norm.bias"]
        self.w1, self.b1 = w[prefix + "linear_fc1.weight"], w[prefix + "linear_fc1.bias"]
        self.w2, self.b2 = w[prefix + "linear_fc2.weight"], w[prefix + "linear_fc2.bias"]

    def forward(self, xs):
        if self.post:
            xs = xs.reshape(-1, self.hidden)
        xs = nn.layer_norm(xs, self.nw, self.nb, 1e-6).reshape(-1, self.hidden)
        return nn.linear(nn.gelu_erf(nn.linear(xs, self.w1, self.b1)), self.w2, self.b2)


class Qwen3VLVisionBlock:
    """model.rs:187-371 (attention + block); MLP = gguf.rs:384-390 fc2(act(fc1(x)))."""

    def __init__(self, vc, w, prefix):
        self.nh = vc["num_heads"]
        self.hd = vc["hidden_size"] // self.nh
        g = lambda n: w[prefix + n]
        self.n1w, self.n1b, self.n2w, self.n2b = g("norm1.weight"), g("norm1.bias"), g("norm2.weight"), g("norm2.bias")
        self.qkv_w, self.qkv_b = g("attn.qkv.weight"), g("attn.qkv.bias")
        self.proj_w, self.proj_b = g("attn.proj.weight"), g("attn.proj.bias")
        self.fc1_w, self.fc1_b = g("mlp.linear_fc1.weight"), g("mlp.linear_fc1.bias")
        self.fc2_w, self.fc2_b = g("mlp.linear_fc2.weight"), g("mlp.linear_fc2.bias")
        self.act = nn.activation(vc.get("hidden_act", "gelu_pytorch_tanh"))
        self.scaling = 1.0 / np.sqrt(np.float64(self.hd))

    def attn(self, xs, cos, sin, cu_seqlens):
        s = xs.shape[0]
        qkv = nn.linear(xs, self.qkv_w, self.qkv_b).reshape(s, 3, self.nh, self.hd)
        q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
        q, k = apply_rotary_pos_emb_vision(q, k, cos, sin)
        q = np.swapaxes(q, 0, 1)[None]
        k = np.swapaxes(k, 0, 1)[None]
        v = np.swapaxes(v, 0, 1)[None]
        outs = []
        for a, b in zip(cu_seqlens[:-1], cu_seqlens[1:]):  # model.rs:258-277: full attention per segment
            a, b = int(a), int(b)
            # per head to bound the (S,S) score tensor
            o = np.concatenate([eager_attention_forward(q[:, h:h + 1, a:b], k[:, h:h + 1, a:b], v[:, h:h + 1, a:b],
                                                        None, None, self.scaling) for h in range(self.nh)], axis=2)
            outs.append(o)
        o = np.concatenate(outs, axis=1).reshape(s, -1)
        return nn.linear(o, self.proj_w, self.proj_b)

    def forward(self, xs, cu_seqlens, cos, sin):
        xs = xs + self.attn(nn.layer_norm(xs, self.n1w, self.n1b, 1e-6), cos, sin, cu_seqlens)
        h = nn.layer_norm(xs, self.n2w, self.n2b, 1e-6)
        xs = xs + nn.linear(self.act(nn.linear(h, self.fc1_w, self.fc1_b)), self.fc2_w, self.fc2_b)
        return xs.astype(F32)


class Qwen3VLVisionModel:
    """model.rs:373-741."""

    def __init__(self, vc, w, prefix="model.visual."):
        self.vc = vc
        self.merge = vc["spatial_merge_size"]
        pw = w[prefix + "patch_embed.proj.weight"]  # (Hv,3,2,16,16)
        self.patch_w = pw.reshape(pw.shape[0], -1)  # flatten(1,4); forward uses its transpose
        self.patch_b = w[prefix + "patch_embed.proj.bias"]
        self.pos_embed = w[prefix + "pos_embed.weight"]
        self.num_grid_per_side = int(np.sqrt(np.float32(vc["num_position_embeddings"])))
        hd = vc["hidden_size"] // vc["num_heads"]
        self.rotary = Qwen2_5VisionRotaryEmbedding(hd // 2)
        self.blocks = [Qwen3VLVisionBlock(vc, w, f"{prefix}blocks.{i}.") for i in range(vc["depth"])]
        self.merger = Qwen3VLVisionPatchMerger(vc, w, prefix + "merger.", False)
        self.ds_idx = list(vc["deepstack_visual_indexes"])
        self.ds_mergers = [Qwen3VLVisionPatchMerger(vc, w, f"{prefix}deepstack_merger_list.{i}.", True)
                           for i in range(len(self.ds_idx))]
        self.trace = None

    def fast_pos_embed_interpolate(self, grid_thw):
        """model.rs:512-639."""
        n = self.num_grid_per_side
        outs = []
        for t, h, w in np.asarray(grid_thw).tolist():
            h_idxs = linspace(0.0, n - 1, h)
            w_idxs = linspace(0.0, n - 1, w)
            hf = h_idxs.astype(np.uint32)  # truncation toward zero
            wf = w_idxs.astype(np.uint32)
            hc = np.clip(hf + 1, 0, n - 1)
            wc = np.clip(wf + 1, 0, n - 1)
            dh = (h_idxs - hf.astype(F32))[:, None]
            dw = (w_idxs - wf.astype(F32))[None, :]
            bh = (hf * n)[:, None]
            bhc = (hc * n)[:, None]
            idx = [bh + wf[None, :], bh + wc[None, :], bhc + wf[None, :], bhc + wc[None, :]]
            wt = [(F32(1) - dh) * (F32(1) - dw), (F32(1) - dh) * dw, dh * (F32(1) - dw), dh * dw]
            pe = None
            for i4 in range(4):
                term = self.pos_embed[idx[i4].reshape(-1).astype(np.int64)].astype(F32) * wt[i4].reshape(-1, 1).astype(F32)
                pe = term if pe is None else (pe + term).astype(F32)
            d = pe.shape[-1]
            pe = np.tile(pe, (t, 1))
            m = self.merge
            pe = pe.reshape(t, h // m, m, w // m, m, d).transpose(0, 1, 3, 2, 4, 5).reshape(-1, d)
            outs.append(pe)
        return np.concatenate(outs, axis=0).astype(F32)

    def rot_pos_emb(self, grid_thw):
        """model.rs:641-690."""
        g = np.asarray(grid_thw)
        table = self.rotary.forward(int(g[:, 1:].max()))
        ids = []
        m = self.merge
        for t, h, w in g.tolist():
            mh, mw = h // m, w // m
            row = (np.arange(mh)[:, None, None, None] * m + np.arange(m)[None, None, :, None])
            col = (np.arange(mw)[None, :, None, None] * m + np.arange(m)[None, None, None, :])
            row = np.broadcast_to(row, (mh, mw, m, m)).reshape(-1)
            col = np.broadcast_to(col, (mh, mw, m, m)).reshape(-1)
            coords = np.stack([row, col], axis=-1)
            if t > 1:
                coords = np.tile(coords, (t, 1))
            ids.append(coords)
        ids = np.concatenate(ids, axis=0)
        return np.concatenate([table[ids[:, 0]], table[ids[:, 1]]], axis=1).astype(F32)

    def forward(self, pixel_values, grid_thw):
        g = np.asarray(grid_thw)
        x = nn.linear(pixel_values, self.patch_w, self.patch_b)  # model.rs:96-103
        x = (x + self.fast_pos_embed_interpolate(g)).astype(F32)
        rot = self.rot_pos_emb(g)
        emb = np.concatenate([rot, rot], axis=-1)
        cos, sin = np.cos(emb).astype(F32), np.sin(emb).astype(F32)
        seg = np.concatenate([np.repeat(h * w, t) for t, h, w in g.tolist()])
        cu = np.concatenate([[0], np.cumsum(seg)]).astype(np.int64)  # model.rs:709-720
        if self.trace is not None:
            self.trace.append(("embed", x.copy()))
        deep = []
        for i, blk in enumerate(self.blocks):
            x = blk.forward(x, cu, cos, sin)
            if self.trace is not None:
                self.trace.append((f"block{i}", x.copy()))
            if i in self.ds_idx:
                deep.append(self.ds_mergers[self.ds_idx.index(i)].forward(x))
        return self.merger.forward(x), deep


# ----------------------------------------------------------------------------- text model + top level
class Qwen3VLTextModel:
    """model.rs:743-835."""

    def __init__(self, tc, w, prefix="model.language_model."):
        self.tc = tc
        self.embed = w[prefix + "embed_tokens.weight"]
        self.layers = [Qwen3DecoderLayer(tc, w, f"{prefix}layers.{i}.") for i in range(tc["num_hidden_layers"])]
        self.norm = w[prefix + "norm.weight"]
        self.rotary = Qwen3VLTextRotaryEmbedding(tc["head_dim"], tc["rope_theta"])
        self.mrope_section = list(tc["rope_scaling"]["mrope_section"])
        self.trace = None

    def forward(self, inputs_embeds, seqlen_offset, position_ids, visual_pos_mask, deepstack):
        b, s, _ = inputs_embeds.shape
        if position_ids is None:
            position_ids = np.broadcast_to(np.arange(seqlen_offset, seqlen_offset + s, dtype=np.int64)[None, None], (3, b, s))
        cos, sin = self.rotary.forward(position_ids, self.mrope_section)
        x = inputs_embeds
        mask = prepare_causal_attention_mask(b, s, 0) if s > 1 else None
        for i, layer in enumerate(self.layers):
            x = layer.forward(x, cos, sin, mask)
            if deepstack is not None and i < len(deepstack):  # model.rs:815-824 mask_index_add
                idx = np.nonzero(visual_pos_mask[0])[0]
                x = x.copy()
                x[0, idx] = (x[0, idx] + deepstack[i]).astype(F32)
            if self.trace is not None:
                self.trace.append(x.copy())
        return nn.rms_norm(x, self.norm, self.tc["rms_norm_eps"])

    def clear_kv_cache(self):
        for l in self.layers:
            l.clear_kv_cache()


def get_rope_index(input_ids, image_grid_thw, cfg, video_grid_thw=None):
    """model.rs:901-1133, mask=None.  Every (t, h, w) row of video_grid_thw becomes t rows of (1, h, w) (model.rs:907-925): each frame
    group has its own <|vision_start|><|video_pad|>... run in the prompt.  Returns (position_ids (3,1,S) int64, rope_delta int)."""
    ids = np.asarray(input_ids).reshape(-1)
    S = ids.shape[0]
    if image_grid_thw is None and video_grid_thw is None:
        pos = np.broadcast_to(np.arange(S, dtype=np.int64)[None, None], (3, 1, S)).copy()
        return pos, 0
    merge = cfg["vision_config"]["spatial_merge_size"]
    img_tok, vid_tok, vs_tok = cfg["image_token_id"], cfg.get("video_token_id"), cfg["vision_start_token_id"]
    g = np.asarray(image_grid_thw) if image_grid_thw is not None else None
    vg = None
    if video_grid_thw is not None:
        vg = [[1, int(h), int(w)] for t, h, w in np.asarray(video_grid_thw).tolist() for _ in range(int(t))]
    vis_next = np.nonzero(ids == vs_tok)[0] + 1  # get_vision_next_indices
    chunks = []
    text_start, text_end, image_index, video_index = 0, 0, 0, 0
    thw = None
    last_max = -1
    for j in vis_next.tolist():
        tok = ids[j] if j < S else -1
        if tok == img_tok:
            thw = g[image_index].tolist()
            image_index += 1
            text_end = j
        if vid_tok is not None and tok == vid_tok:
            thw = vg[video_index]
            video_index += 1
            text_end = j
        if thw is None:
            continue
        gt, gh, gw = thw[0], thw[1] // merge, thw[2] // merge
        text_len = text_end - text_start
        start = last_max + 1 if chunks else 0
        chunks.append(np.broadcast_to(np.arange(start, start + text_len, dtype=np.int64)[None], (3, text_len)))
        base = start + text_len
        t_idx = np.broadcast_to(np.arange(base, base + gt)[:, None], (gt, gh * gw)).reshape(-1)
        h_idx = np.broadcast_to(np.arange(base, base + gh)[None, :, None], (gt, gh, gw)).reshape(-1)
        w_idx = np.broadcast_to(np.arange(base, base + gw)[None, None, :], (gt, gh, gw)).reshape(-1)
        blk = np.stack([t_idx, h_idx, w_idx], axis=0).astype(np.int64)
        chunks.append(blk)
        last_max = int(blk.max())  # "max of the last pushed chunk"
        text_start = text_end + gt * gh * gw
    if text_start < S:
        start = (int(chunks[-1].max()) + 1) if chunks else 0
        text_len = S - text_start
        chunks.append(np.broadcast_to(np.arange(start, start + text_len, dtype=np.int64)[None], (3, text_len)))
    pos = np.concatenate(chunks, axis=1).reshape(3, 1, -1)
    delta = int(pos.max()) + 1 - S
    return pos, delta


class Qwen3VLModel:
    """model.rs:837-1324 incl. `impl InferenceModel` (image and video tensors; decoding a video file into frames is the processor's job)."""

    def __init__(self, cfg, w, eos_ids=()):
        self.cfg = cfg
        self.visual = Qwen3VLVisionModel(cfg["vision_config"], w)
        tc = dict(cfg["text_config"])
        self.text = Qwen3VLTextModel(tc, w)
        self.lm_head = self.text.embed if cfg.get("tie_word_embeddings", False) else w["lm_head.weight"]
        self.rope_deltas = None
        self._stop = list(eos_ids)

    def forward(self, input_ids, pixel_values=None, image_grid_thw=None, seqlen_offset=0, pixel_values_video=None, video_grid_thw=None):
        """model.rs:1135-1290."""
        ids = np.asarray(input_ids).reshape(1, -1)
        x = nn.embedding(ids, self.text.embed)
        image_mask = video_mask = None
        deep_img = deep_vid = None
        if pixel_values is not None and image_grid_thw is not None:
            emb, deep_img = self.visual.forward(pixel_values, image_grid_thw)
            image_mask = (ids == self.cfg["image_token_id"])
            n_tok = int(image_mask.sum())
            if n_tok != emb.shape[0]:  # model.rs:1158-1164
                raise ValueError(f"n_image_token num: {n_tok} not equal to image_embed len: {emb.shape[0]}")
            x = x.copy()
            x[0, np.nonzero(image_mask[0])[0]] = emb  # masked_scatter_dim0
        if pixel_values_video is not None and video_grid_thw is not None:
            emb, deep_vid = self.visual.forward(pixel_values_video, video_grid_thw)
            video_mask = (ids == self.cfg["video_token_id"])
            n_tok = int(video_mask.sum())
            if n_tok != emb.shape[0]:  # model.rs:1176-1183 (the same message)
                raise ValueError(f"n_image_token num: {n_tok} not equal to image_embed len: {emb.shape[0]}")
            x = x.copy()
            x[0, np.nonzero(video_mask[0])[0]] = emb
        mask, deep = None, None
        if image_mask is not None and video_mask is not None:  # model.rs:1189-1218: joint embedding in visual-position order
            mask = image_mask | video_mask
            vis = np.nonzero(mask[0])[0]
            img_joint = np.nonzero(image_mask[0][vis])[0]
            vid_joint = np.nonzero(video_mask[0][vis])[0]
            deep = []
            for di, dv in zip(deep_img, deep_vid):
                joint = np.zeros((vis.shape[0], di.shape[-1]), F32)
                joint[img_joint] += di
                joint[vid_joint] += dv
                deep.append(joint)
        elif image_mask is not None:
            mask, deep = image_mask, deep_img
        elif video_mask is not None:
            mask, deep = video_mask, deep_vid
        if self.rope_deltas is None:
            pos, delta = get_rope_index(ids, image_grid_thw, self.cfg, video_grid_thw)
            self.rope_deltas = delta
        else:  # model.rs:1235-1264
            s = ids.shape[1]
            pos = np.broadcast_to((np.arange(s, dtype=np.int64) + seqlen_offset + self.rope_deltas)[None, None], (3, 1, s))
        out = self.text.forward(x, seqlen_offset, pos, mask, deep)
        s = out.shape[1]
        return nn.linear(out[:, s - 1:s, :], self.lm_head)

    def forward_initial(self, input_ids, seqlen_offset, data):
        if data is None or len(data) != 5:  # model.rs:1292-1296
            raise ValueError("Qwen3VL process data error, must have pixel_values, image_grid_thw, "
                             "pixel_values_video, video_grid_thw, cache_position")
        return self.forward(input_ids, data[0], data[1], seqlen_offset, data[2], data[3])

    def forward_step(self, input_ids, seqlen_offset):
        return self.forward(input_ids, None, None, seqlen_offset)

    def clear_cache(self):
        self.rope_deltas = None
        self.text.clear_kv_cache()

    def stop_token_ids(self):
        return list(self._stop)
