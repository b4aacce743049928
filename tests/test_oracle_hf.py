"""Cross-check of the oracle against HF transformers where the reference has no quirk (text decoder, ViT,
pos-embed interpolation, 2-D RoPE, deepstack, M-RoPE, get_rope_index, slaney mel bank).  CPU only."""
import numpy as np
import pytest

from aha_b200 import synth

torch = pytest.importorskip("torch")
transformers = pytest.importorskip("transformers")


def test_qwen3_matches_hf():
    from transformers import Qwen3Config, Qwen3ForCausalLM
    from oracle.qwen3 import Qwen3Model
    cfg = synth.get_config("qwen3", "tiny")
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


def test_mel_filter_bank_matches_hf():
    from transformers.audio_utils import mel_filter_bank as hfmel
    from oracle.audio import mel_filter_bank
    h = hfmel(201, 128, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney")
    assert np.abs(h - mel_filter_bank(201, 128, 0.0, 8000.0, 16000)).max() < 1e-6
