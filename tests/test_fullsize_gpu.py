"""BASELINE.json's full sizes, checked through size-independent properties (the oracle would need minutes per
forward at these shapes):
  * cache consistency: forward_initial(ids[:S]) followed by forward_step(ids[S]) must give the logits of
    forward_initial(ids[:S+1]) -- ties the prefill kernels (tcgen05 GEMM, flash attention, paged KV write) to the
    decode kernels (fused step / per-op) on the same weights;
  * the two decode implementations agree; greedy decode is deterministic across requests;
  * GPU log-mel of 30 s of audio equals the oracle's (cheap on the CPU), 3000 frames -> 390 audio tokens;
  * 1088x1920 image -> 8160 patches -> 2040 image tokens, rope_delta = -1980 (closed forms from the source)."""
import numpy as np
import pytest

from conftest import TOL
from aha_b200 import B200Model, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vl2():
    cfg = synth.get_config("qwen3vl", "vl2")
    w = synth.make_weights("qwen3vl", cfg, 0)
    m = B200Model("qwen3vl", cfg, w, eos_ids=[], max_ctx=4096, max_prefill=4096, max_patches=8192)
    m1 = B200Model("qwen3vl", cfg, w, eos_ids=[], max_ctx=4096, max_prefill=4096, max_patches=8192, decode_impl=1)
    del w
    yield cfg, m, m1
    m.close(); m1.close()


def test_vl2_1080p_cache_consistency_and_closed_forms(vl2):
    cfg, m, m1 = vl2
    img = synth.synth_image(1088, 1920, 1)
    pv, grid = m.image_patchify(img)
    assert pv.shape == (8160, 1536) and grid.tolist() == [[1, 68, 120]]
    ids = synth.vl_prompt_ids(cfg, grid, 512)
    assert int((ids == cfg["image_token_id"]).sum()) == 2040 and len(ids) == 2554
    data = [pv, grid, None, None, None]
    S = len(ids) - 1
    # request A: prefill S tokens, then one decode step with token S (fused kernel)
    m.forward_initial(ids[:S], 0, data, want_logits=False)
    assert int(m.debug_read("rope_delta", 0, 1)[0]) == -1980
    a = m.forward_step(ids[S:S + 1], S)[0, 0]
    # request B: prefill all S+1 tokens
    m.clear_cache()
    b = m.forward_initial(ids, 0, data)[0, 0]
    err = float(np.abs(a - b).max())
    assert err <= TOL, err
    # per-op decode implementation on the same request
    m1.forward_initial(ids[:S], 0, data, want_logits=False)
    c = m1.forward_step(ids[S:S + 1], S)[0, 0]
    assert float(np.abs(c - b).max()) <= TOL
    print(f"\nVL2 1080p: |prefill(S)+step - prefill(S+1)| = {err:.2e} (fused), {float(np.abs(c - b).max()):.2e} (per-op); logit std {b.std():.3f}")


def test_vl2_greedy_decode_is_deterministic_and_impls_agree(vl2):
    cfg, m, m1 = vl2
    ids = synth.synth_text_ids(600, 151000, 4)
    m.clear_cache(); m1.clear_cache()
    t0, _ = m.generate(ids, [None] * 5, max_tokens=24)
    t1, _ = m.generate(ids, [None] * 5, max_tokens=24)
    t2, _ = m1.generate(ids, [None] * 5, max_tokens=24)
    assert t0 == t1
    assert t0 == t2


def test_q06_config2_2k_context():
    """config 2: Qwen3-0.6B shape, 1920-token prompt + decode to ctx 2048 (fp32 KV)."""
    cfg = synth.get_config("qwen3", "q0.6")
    w = synth.make_weights("qwen3", cfg, 0)
    m = B200Model("qwen3", cfg, w, eos_ids=[], max_ctx=2048, max_prefill=2048)
    del w
    try:
        ids = synth.synth_text_ids(1921, 151000, 9)
        m.forward_initial(ids[:1920], 0, want_logits=False)
        a = m.forward_step(ids[1920:1921], 1920)[0, 0]
        m.clear_cache()
        b = m.forward_initial(ids, 0)[0, 0]
        assert float(np.abs(a - b).max()) <= TOL
        m.clear_cache()
        toks, usage = m.generate(ids[:1920], max_tokens=128)       # runs to ctx 2048 exactly
        assert len(toks) == 128 and usage["prompt_tokens"] == 1920
        with pytest.raises(Exception, match="max_ctx"):
            m.generate(ids[:1921], max_tokens=128)
    finally:
        m.close()


def test_asr06_config4_30s_audio():
    from oracle.audio import WhisperFeatureExtractor, get_feat_extract_output_lengths
    cfg = synth.get_config("qwen3_asr", "asr0.6")
    w = synth.make_weights("qwen3_asr", cfg, 0)
    m = B200Model("qwen3_asr", cfg, w, eos_ids=[], max_ctx=1024, max_frames=3000)
    del w
    try:
        wave = synth.synth_audio(30.0)
        mel = m.mel_spectrogram(wave)
        assert mel.shape == (128, 3000)
        want = WhisperFeatureExtractor().call(wave[None], 16000)[0]
        assert float(np.abs(mel - want).max()) <= 1e-3
        n_tok = get_feat_extract_output_lengths(3000)
        assert n_tok == 390
        ids = synth.asr_prompt_ids(cfg, n_tok, n_text=9)
        S = len(ids) - 1
        m.forward_initial(ids[:S], 0, [mel], want_logits=False)
        assert m.debug_read("audio_embeds", 0, 390 * 1024).size == 390 * 1024
        a = m.forward_step(ids[S:S + 1], S)[0, 0]
        m.clear_cache()
        b = m.forward_initial(ids, 0, [mel])[0, 0]
        assert float(np.abs(a - b).max()) <= TOL
    finally:
        m.close()
