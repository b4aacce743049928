"""Device sampler, streaming and the ASR loop against oracle/sample.py and oracle/generate.py.

 * the sampler kernel on GIVEN logits (aha_b200_debug_sample) for every Sampling the reference can select -- All, TopP, TopK,
   TopKThenTopP -- plus the repeat penalty: same ChaCha12 stream, same blocked sums, same token as the oracle;
 * generate() / generate_stream() / asr_generate() on the tiny models against the reference loop restated in the oracle
   (first token never EOS-checked for generate_generic, always for the ASR loop; one RNG stream across ASR chunks)."""
import numpy as np
import pytest

from conftest import make_model, make_oracle
from oracle import sample as S

pytestmark = pytest.mark.gpu

MODES = {"all": (0.8, None, None), "topp": (0.7, 0.8, None), "topk": (0.9, None, 20), "topk_topp": (0.6, 0.9, 20),
         "topk_topp_small_p": (1.3, 0.3, 50), "topk_1": (0.7, None, 1), "topp_tiny": (0.7, 0.05, None)}


@pytest.fixture(scope="module")
def tiny():
    cfg, w, m = make_model("qwen3", "tiny", max_ctx=512)
    yield cfg, w, m
    m.close()


@pytest.mark.parametrize("mode", sorted(MODES))
def test_sampler_kernel_matches_the_oracle_on_given_logits(tiny, mode):
    cfg, w, m = tiny
    t, p, k = MODES[mode]
    V = cfg["vocab_size"]
    rng = np.random.default_rng(5)
    oracle_rng = S.StdRng(1234)
    mism = 0
    for trial in range(12):
        lg = (rng.standard_normal(V) * rng.uniform(0.5, 4.0)).astype(np.float32)
        ctx = rng.integers(0, V, size=rng.integers(0, 90)).astype(np.uint32)
        pen = 1.0 if trial % 2 else 1.3
        want_l = S.use_repeat_penalty(pen, 64, lg, list(ctx))
        want = S.sample(want_l, t, p, k, oracle_rng, trial)
        got = m.debug_sample(lg, context=ctx, draw_index=trial, temperature=t, top_p=p, top_k=k, repeat_penalty=pen, seed=1234)
        mism += int(got != want)
    assert mism == 0, f"{mism} of 12 sampled tokens differ from the oracle"


def test_repeat_penalty_with_argmax_on_given_logits(tiny):
    cfg, w, m = tiny
    V = cfg["vocab_size"]
    rng = np.random.default_rng(6)
    for trial in range(6):
        lg = rng.standard_normal(V).astype(np.float32)
        ctx = np.concatenate([np.argsort(-lg)[:3], rng.integers(0, V, 70)]).astype(np.uint32)   # the three best tokens are in the context
        rng.shuffle(ctx)
        want = int(np.argmax(S.use_repeat_penalty(1.7, 64, lg, list(ctx))))
        assert m.debug_sample(lg, context=ctx, temperature=0.0, repeat_penalty=1.7, repeat_last_n=64) == want


def _oracle_generate(o, ids, sampler, max_tokens, eos, eos_on_first=False, data=None):
    gen = []
    logits = o.forward_initial(ids.reshape(1, -1), 0, data) if data is not None else o.forward_initial(ids.reshape(1, -1), 0)
    tok = sampler.sample_and_push(logits, gen)
    off = len(ids)
    if eos_on_first and tok in eos:
        o.clear_cache()
        return gen
    for _ in range(1, max_tokens):
        logits = o.forward_step(np.array([[tok]]), off)
        off += 1
        tok = sampler.sample_and_push(logits, gen)
        if tok in eos:
            break
    o.clear_cache()
    return gen


@pytest.mark.parametrize("mode", ["all", "topk_topp", "topp"])
def test_generate_with_sampling_matches_the_oracle_loop(tiny, mode):
    cfg, w, m = tiny
    t, p, k = MODES[mode]
    o = make_oracle("qwen3", cfg, w)
    ids = np.random.default_rng(3).integers(0, 1000, 40).astype(np.uint32)
    want = _oracle_generate(o, ids, S.Sampler(t, p, k, 1.1, 16, seed=77), 24, [cfg["eos_token_id"]])
    got, usage = m.generate(ids, max_tokens=24, temperature=t, top_p=p, top_k=k, repeat_penalty=1.1, repeat_last_n=16, seed=77)
    assert got == want
    assert usage["completion_tokens"] == len(got)


def test_generate_stream_delivers_the_same_tokens_and_can_be_stopped(tiny):
    cfg, w, m = tiny
    ids = np.random.default_rng(4).integers(0, 1000, 33).astype(np.uint32)
    want, _ = m.generate(ids, max_tokens=21, temperature=0.7, top_k=20, top_p=0.8, seed=5)
    seen = []
    usage = m.generate_stream(ids, lambda t, i: seen.append((i, t)) and False, max_tokens=21, temperature=0.7, top_k=20, top_p=0.8, seed=5)
    assert [t for _, t in seen] == want and [i for i, _ in seen] == list(range(len(want)))
    assert usage["completion_tokens"] == len(want)
    seen2 = []
    m.generate_stream(ids, lambda t, i: (seen2.append(t), i == 4)[1], max_tokens=21, temperature=0.7, top_k=20, top_p=0.8, seed=5)
    assert seen2 == want[:5]                      # a truthy return ends the request after that token
    greedy, _ = m.generate(ids, max_tokens=9)
    s3 = []
    m.generate_stream(ids, lambda t, i: s3.append(t) and False, max_tokens=9)
    assert s3 == greedy


def test_max_tokens_zero_yields_one_token(tiny):
    cfg, w, m = tiny
    ids = np.arange(5, 25, dtype=np.uint32)
    a, _ = m.generate(ids, max_tokens=0)
    b, _ = m.generate(ids, max_tokens=1)
    assert len(a) == 1 and a == b


def test_asr_generate_loops_over_chunks_with_one_sampler():
    from oracle.audio import WhisperFeatureExtractor, get_feat_extract_output_lengths
    from aha_b200 import synth
    cfg, w, m = make_model("qwen3_asr", "tiny", max_ctx=512, max_frames=400)
    o = make_oracle("qwen3_asr", cfg, w)
    eos = [cfg["thinker_config"]["text_config"]["eos_token_id"], 7]
    m.close()
    from aha_b200 import B200Model
    m = B200Model("qwen3_asr", cfg, w, eos_ids=eos, max_ctx=512, max_frames=400)
    try:
        chunks = []
        for sec, seed in ((2.5, 2), (1.3, 3)):
            mel = WhisperFeatureExtractor().call(synth.synth_audio(sec, seed=seed)[None], 16000)[0]
            chunks.append((synth.asr_prompt_ids(cfg, get_feat_extract_output_lengths(mel.shape[1])), mel))
        sampler = S.Sampler(0.8, 0.9, None, None, None, seed=34562)       # get_logit_processor(Some(temperature), top_p, None, seed)
        want = []
        for ids, mel in chunks:
            want += _oracle_generate_asr(o, ids, mel, sampler, 12, eos)
        got, usage = m.asr_generate(chunks, max_tokens=12, temperature=0.8, top_p=0.9, seed=34562)
        assert got == want
        assert usage["prompt_tokens"] == sum(len(i) for i, _ in chunks)
        streamed = []
        got2, _ = m.asr_generate(chunks, max_tokens=12, temperature=0.8, top_p=0.9, seed=34562, on_token=lambda t, i: streamed.append(t) and False)
        assert got2 == want and streamed == want
    finally:
        m.close()


def _oracle_generate_asr(o, ids, mel, sampler, sample_len, eos):
    """qwen3_asr/generate.rs:148-173: every token (the first too) is EOS-checked, no repeat penalty, cache cleared per chunk."""
    gen = []
    x, off, feat = ids.reshape(1, -1), 0, mel
    for _ in range(sample_len):
        logits = o.forward(x, off, feat)
        tok = sampler.sample_and_push(logits, gen, penalise=False)
        if tok in eos:
            break
        off += x.shape[1]
        x, feat = np.array([[tok]]), None
    o.clear_cache()
    return gen
