"""KV reuse across requests and prefill continuation (new design -- the reference clears its cache after every
request, common/generate.rs:147, and rejects a multi-token call against a non-empty cache, qwen3/model.rs:164-175).

The reference's answer is the one a FRESH request gives, so that is the oracle here: whatever the cache supplies, the logits
must match the oracle's full-prompt forward_initial (<= 1e-3) and the generated ids must equal the oracle's generate_generic."""
import numpy as np
import pytest

from conftest import TOL, make_model, make_oracle, top2_gap

pytestmark = pytest.mark.gpu


def _ids(n, vocab, seed):
    from aha_b200 import synth
    return synth.synth_text_ids(n, vocab - 8, seed)


def _oracle_generate(o, ids, data, n, **samp):
    from oracle.generate import GenerationContext, generate_generic
    o.clear_cache()
    if not samp:
        ctx = GenerationContext(temperature=0.0, initial_seq_len=len(ids), max_tokens=n)
        return generate_generic(o, np.asarray(ids).reshape(1, -1), data, ctx)[0]
    from oracle.sample import Sampler
    s = Sampler(samp.get("temperature"), samp.get("top_p"), samp.get("top_k"), samp.get("repeat_penalty", 1.0), samp.get("repeat_last_n", 64),
                seed=samp.get("seed", 299792458))
    toks = []
    eos = o.stop_token_ids()
    t = s.sample_and_push(o.forward_initial(np.asarray(ids).reshape(1, -1), 0, data), toks)
    for i in range(1, n):
        t = s.sample_and_push(o.forward_step(np.array([[t]]), len(ids) + i - 1), toks)
        if t in eos:
            break
    o.clear_cache()
    return toks


@pytest.fixture(scope="module")
def q3():
    cfg, w, m = make_model("qwen3", "tiny", max_ctx=1024)
    yield cfg, w, m, make_oracle("qwen3", cfg, w)
    m.close()


@pytest.mark.parametrize("attn_impl,gemm_impl", [(0, 0), (1, 1), (2, 0)])
@pytest.mark.parametrize("cuts", [[1], [64], [37, 38], [5, 70, 199], [128, 256]])
def test_forward_extend_matches_full_prefill(cuts, attn_impl, gemm_impl):
    """forward_initial(ids[:a]) + forward_extend(ids[a:b], a) + ... == the oracle's forward_initial(ids): every prefill attention
    kernel (tcgen05, fp32 SIMT twin, mma.sync) with the (S, offset + S) causal mask over the paged cache, then decode on top."""
    cfg, w, m = make_model("qwen3", "tiny", max_ctx=512, attn_impl=attn_impl, gemm_impl=gemm_impl)
    o = make_oracle("qwen3", cfg, w)
    S = 300
    ids = _ids(S + 6, cfg["vocab_size"], 41)
    want = o.forward_initial(ids[:S].reshape(1, -1), 0)[0, 0]
    bounds = [0] + cuts + [S]
    got = m.forward_initial(ids[:bounds[1]], 0)[0, 0]
    for a, b in zip(bounds[1:-1], bounds[2:]):
        got = m.forward_extend(ids[a:b], a)[0, 0]
    err = np.abs(got - want).max()
    assert err <= (2e-5 if attn_impl == 1 else TOL), err
    if top2_gap(want) > 10 * err:
        assert m.last_argmax == int(np.argmax(want))
    for i in range(6):   # the cache the chunks built is the cache a single prefill builds: decode steps agree with the oracle
        got = m.forward_step(ids[S + i:S + i + 1], S + i)[0, 0]
        want = o.forward_step(ids[S + i:S + i + 1].reshape(1, 1), S + i)[0, 0]
        assert np.abs(got - want).max() <= TOL
    m.close()


def test_forward_extend_errors(q3):
    from aha_b200 import B200Error
    cfg, w, m, o = q3
    m.clear_cache()
    with pytest.raises(B200Error, match="beyond the tokens in the cache"):
        m.forward_extend(_ids(4, cfg["vocab_size"], 1), 100)
    m.forward_initial(_ids(10, cfg["vocab_size"], 1), 0)
    with pytest.raises(B200Error, match="not supported"):          # the trait call keeps the reference's behaviour
        m.forward_step(_ids(4, cfg["vocab_size"], 2), 10)
    m.clear_cache()


def test_prompt_longer_than_max_prefill_runs_in_chunks():
    """max_prefill bounds the activation workspace, not the prompt: a longer prompt is cut into continuations."""
    cfg, w, m = make_model("qwen3", "tiny", max_ctx=512, max_prefill=96)
    o = make_oracle("qwen3", cfg, w)
    ids = _ids(250, cfg["vocab_size"], 9)
    want = o.forward_initial(ids.reshape(1, -1), 0)[0, 0]
    got = m.forward_initial(ids, 0)[0, 0]
    assert np.abs(got - want).max() <= TOL
    o.clear_cache()
    toks, _ = m.generate(ids, max_tokens=12)
    assert toks == _oracle_generate(o, ids, None, 12)
    m.close()


def test_multi_turn_reuses_the_conversation(q3):
    """Turn 2's prompt = turn 1's prompt + its answer + new text: only the new text is prefilled; ids equal the oracle's fresh runs."""
    cfg, w, m, o = q3
    V = cfg["vocab_size"]
    m.clear_cache()
    p1 = _ids(90, V, 21)
    a1, _ = m.generate(p1, max_tokens=17, reuse_prefix=True)
    assert m.last_prefix_hit() == 0
    assert a1 == _oracle_generate(o, p1, None, 17)
    p2 = np.concatenate([p1, np.asarray(a1, np.uint32), _ids(33, V, 22)])
    a2, u2 = m.generate(p2, max_tokens=20, reuse_prefix=True)
    assert m.last_prefix_hit() == len(p1) + len(a1) - 1          # the last generated token was never fed back: its K/V is not there
    assert a2 == _oracle_generate(o, p2, None, 20)
    assert u2["prompt_tokens"] == len(p2)
    # turn 3 diverges inside turn 2's prompt: the common prefix is reused, the rest overwritten
    p3 = np.concatenate([p2[:120], _ids(40, V, 23)])
    a3, _ = m.generate(p3, max_tokens=9, reuse_prefix=True)
    assert m.last_prefix_hit() == 120
    assert a3 == _oracle_generate(o, p3, None, 9)
    # the same prompt again: everything but the last token comes from the cache
    a3b, _ = m.generate(p3, max_tokens=9, reuse_prefix=True)
    assert m.last_prefix_hit() == len(p3) - 1 and a3b == a3
    # a request without the flag starts from clear_cache() like the reference, and leaves nothing behind
    a3c, _ = m.generate(p3, max_tokens=9)
    assert m.last_prefix_hit() == 0 and a3c == a3
    m.generate(p3, max_tokens=3, reuse_prefix=True)
    assert m.last_prefix_hit() == 0


def test_reuse_with_the_device_sampler(q3):
    """Sampling state (RNG stream, repeat-penalty history) belongs to the request, not to the cache."""
    cfg, w, m, o = q3
    V = cfg["vocab_size"]
    m.clear_cache()
    samp = dict(temperature=0.8, top_p=0.9, top_k=30, repeat_penalty=1.15, repeat_last_n=16, seed=7)
    p1 = _ids(50, V, 31)
    a1, _ = m.generate(p1, max_tokens=10, reuse_prefix=True, **samp)
    assert a1 == _oracle_generate(o, p1, None, 10, **samp)
    p2 = np.concatenate([p1, np.asarray(a1, np.uint32), _ids(12, V, 32)])
    a2, _ = m.generate(p2, max_tokens=14, reuse_prefix=True, **samp)
    assert m.last_prefix_hit() > 0
    assert a2 == _oracle_generate(o, p2, None, 14, **samp)
    m.clear_cache()


def test_direct_trait_calls_invalidate_the_prefix_cache(q3):
    cfg, w, m, o = q3
    V = cfg["vocab_size"]
    m.clear_cache()
    p = _ids(40, V, 51)
    m.generate(p, max_tokens=4, reuse_prefix=True)
    m.forward_initial(_ids(40, V, 52), 0)           # overwrites the pages behind the cache's back
    toks, _ = m.generate(p, max_tokens=4, reuse_prefix=True)
    assert m.last_prefix_hit() == 0
    assert toks == _oracle_generate(o, p, None, 4)
    m.generate(p, max_tokens=4, reuse_prefix=True)
    m.clear_cache()
    m.generate(p, max_tokens=4, reuse_prefix=True)
    assert m.last_prefix_hit() == 0
    m.clear_cache()


@pytest.mark.parametrize("decode_impl", [1, 3])
def test_streaming_and_decode_paths_with_reuse(decode_impl):
    cfg, w, m = make_model("qwen3", "tiny", max_ctx=512, decode_impl=decode_impl)
    o = make_oracle("qwen3", cfg, w)
    V = cfg["vocab_size"]
    p1 = _ids(70, V, 61)
    got = []
    m.generate_stream(p1, lambda t, i: got.append(t) and False, max_tokens=11, reuse_prefix=True)
    assert got == _oracle_generate(o, p1, None, 11)
    p2 = np.concatenate([p1, np.asarray(got, np.uint32), _ids(5, V, 62)])
    got2 = []
    m.generate_stream(p2, lambda t, i: got2.append(t) and False, max_tokens=11, reuse_prefix=True)
    assert m.last_prefix_hit() == len(p1) + len(got) - 1
    assert got2 == _oracle_generate(o, p2, None, 11)
    m.close()


def test_vl_multi_turn_skips_the_vision_tower():
    """Second turn about the same image: the image rows, their M-RoPE positions and rope_delta all come from the cache; a different
    image under the same token ids is a miss (the fingerprint of pixel_values decides)."""
    from aha_b200 import synth
    from oracle.qwen3vl import process_image
    cfg, w, m = make_model("qwen3vl", "tiny", max_ctx=1024, max_patches=1024)
    o = make_oracle("qwen3vl", cfg, w)
    V = cfg["text_config"]["vocab_size"]
    pv, grid = process_image(synth.synth_image(256, 320, 1))
    data = [pv, grid, None, None, None]
    p1 = np.concatenate([_ids(3, V, 70), synth.vl_prompt_ids(cfg, grid, 24)])   # system text, the image, the question
    a1, u1 = m.generate(p1, data, max_tokens=8, reuse_prefix=True)
    assert a1 == _oracle_generate(o, p1, data, 8)
    assert u1["vision_secs"] > 0
    p2 = np.concatenate([p1, np.asarray(a1, np.uint32), _ids(19, V, 71)])
    a2, u2 = m.generate(p2, data, max_tokens=8, reuse_prefix=True)
    assert m.last_prefix_hit() == len(p1) + len(a1) - 1
    assert u2["vision_secs"] == 0
    assert a2 == _oracle_generate(o, p2, data, 8)
    # same ids, another image: nothing may be reused
    pv_b, grid_b = process_image(synth.synth_image(256, 320, 2))
    data_b = [pv_b, grid_b, None, None, None]
    a3, u3 = m.generate(p2, data_b, max_tokens=8, reuse_prefix=True)
    assert m.last_prefix_hit() == 0 and u3["vision_secs"] > 0
    assert a3 == _oracle_generate(o, p2, data_b, 8)
    # a prompt that diverges BEFORE the image cannot reuse the image rows: they would sit at other positions
    p4 = p2.copy()
    p4[1] = (p4[1] + 1) % 100 + 10
    a4, _ = m.generate(p4, data_b, max_tokens=4, reuse_prefix=True)
    assert m.last_prefix_hit() == 0
    assert a4 == _oracle_generate(o, p4, data_b, 4)
    m.close()
