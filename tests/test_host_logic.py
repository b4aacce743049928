"""Host-side logic of the library that needs no GPU: the C++ M-RoPE index builder behind aha_b200_forward_initial
(aha_b200_rope_index) against the oracle restatement of Qwen3VLModel::get_rope_index
(/root/reference/src/models/qwen3vl/model.rs:901-1133), which tests/test_oracle_hf.py in turn pins against HF."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from aha_b200 import B200Error, rope_index, synth
from oracle.qwen3vl import get_rope_index

CFG = synth.get_config("qwen3vl", "tiny")


def build_prompt(grids, texts, seed=30):
    ids = []
    for k, n in enumerate(texts):
        ids += synth.synth_text_ids(n, 1000, seed + k).tolist()
        if k < len(grids):
            ids += synth.vl_prompt_ids(CFG, [grids[k]], 0).tolist()
    return np.asarray(ids, dtype=np.uint32)


@pytest.mark.parametrize("grids,texts", [
    ([[1, 4, 6]], [5, 7]),
    ([[1, 4, 6], [1, 8, 2]], [3, 4, 9]),
    ([[1, 2, 2], [1, 6, 10], [1, 4, 4]], [0, 1, 0, 6]),
    ([[1, 68, 120]], [0, 512]),
])
def test_rope_index_matches_oracle(grids, texts):
    ids = build_prompt(grids, texts)
    pos, delta = rope_index(ids, grids, CFG)
    want, want_delta = get_rope_index(ids.astype(np.int64), np.asarray(grids), CFG)
    assert pos.shape == (3, ids.size)
    assert (pos == want[:, 0]).all()
    assert delta == want_delta


def test_rope_index_1080p_known_answer():  # rope_deltas = max(34, 60) - 2040, whatever text surrounds the image
    for texts in ([0, 512], [7, 100], [1, 0]):
        _, delta = rope_index(build_prompt([[1, 68, 120]], texts), [[1, 68, 120]], CFG)
        assert delta == -1980


def test_rope_index_text_only_is_arange():
    ids = synth.synth_text_ids(17, 1000, 3)
    pos, delta = rope_index(ids, None, CFG)
    assert delta == 0 and (pos == np.arange(17)[None]).all()


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.integers(1, 6), st.integers(1, 6)), min_size=1, max_size=4), st.data())
def test_rope_index_random_prompts(shapes, data):
    grids = [[1, 2 * h, 2 * w] for h, w in shapes]
    texts = [data.draw(st.integers(0, 9)) for _ in range(len(grids) + 1)]
    ids = build_prompt(grids, texts, seed=data.draw(st.integers(0, 1000)))
    pos, delta = rope_index(ids, grids, CFG)
    want, want_delta = get_rope_index(ids.astype(np.int64), np.asarray(grids), CFG)
    assert (pos == want[:, 0]).all() and delta == want_delta
    assert (pos[:, 1:].max(0) >= pos[:, :-1].min(0)).all()          # positions never jump backwards past a whole chunk
    assert delta == int(pos.max()) + 1 - ids.size


@pytest.mark.parametrize("grids,vgrids,order", [
    ([], [[3, 4, 6]], "v"),                       # video only: three frame groups, each its own (1, h, w) run
    ([[1, 4, 6]], [[2, 8, 2]], "iv"),             # image then video
    ([[1, 2, 2]], [[2, 4, 4], [1, 6, 2]], "vi"),  # two videos then an image
])
def test_rope_index_video_branch_matches_oracle(grids, vgrids, order):
    parts = {"i": [synth.vl_prompt_ids(CFG, [g], 0) for g in grids], "v": [synth.vl_video_prompt_ids(CFG, [g]) for g in vgrids]}
    ids = [synth.synth_text_ids(4, 1000, 1)]
    for k in order:
        ids += parts[k] + [synth.synth_text_ids(2, 1000, 2)]
    ids = np.concatenate(ids).astype(np.uint32)
    pos, delta = rope_index(ids, grids or None, CFG, vgrids)
    want, want_delta = get_rope_index(ids.astype(np.int64), np.asarray(grids) if grids else None, CFG, np.asarray(vgrids))
    assert (pos == want[:, 0]).all() and delta == want_delta
    assert delta == int(pos.max()) + 1 - ids.size
    # a frame group is positioned as a (1, h, w) grid: its t row is constant
    first = int(np.nonzero(ids == CFG["video_token_id"])[0][0])
    n0 = vgrids[0][1] * vgrids[0][2] // 4
    assert (pos[0, first:first + n0] == pos[0, first]).all()


def test_rope_index_errors_are_reported_not_thrown():
    ids = build_prompt([[1, 4, 6], [1, 4, 6]], [2, 2, 2])
    with pytest.raises(B200Error, match="more image placeholders"):
        rope_index(ids, [[1, 4, 6]], CFG)                           # two placeholders, one grid row
    with pytest.raises(B200Error, match="does not cover|exceed"):
        rope_index(ids, [[1, 4, 6], [1, 8, 8]], CFG)                # grid larger than the placeholder run
    tail = np.append(synth.synth_text_ids(4, 1000, 1), np.uint32(CFG["vision_start_token_id"]))
    with pytest.raises(B200Error, match="vision_start"):
        rope_index(tail, [[1, 4, 6]], CFG)
    vid = synth.vl_video_prompt_ids(CFG, [[2, 4, 4]])
    with pytest.raises(B200Error, match="more video placeholder runs"):
        rope_index(vid, None, CFG, [[1, 4, 4]])                      # two frame groups in the prompt, one in video_grid_thw


# ---- KV reuse across requests: the host rule behind AHA_GEN_REUSE_PREFIX (include/aha_b200.h: aha_b200_prefix_match) ----
def _prefix_rule(cached, ids, mm_tokens, same_mm):
    """Plain restatement: common prefix, cut to n - 1; placeholders of either sequence past it (or other tensors) => 0."""
    if len(ids) == 0 or len(cached) == 0 or not same_mm:
        return 0
    lcp = 0
    while lcp < min(len(ids), len(cached)) and cached[lcp] == ids[lcp]:
        lcp += 1
    if any(t in mm_tokens for t in ids[lcp:]) or any(t in mm_tokens for t in cached[lcp:]):
        return 0
    return min(lcp, len(ids) - 1)


def test_prefix_match_known_answers():
    from aha_b200.inference import prefix_match
    IMG = 9
    assert prefix_match([1, 2, 3, 4], [1, 2, 3, 4, 5, 6]) == 4                 # multi-turn: the whole conversation so far
    assert prefix_match([1, 2, 3, 4], [1, 2, 3, 4]) == 3                       # the last prompt token is always run
    assert prefix_match([1, 2, 3, 4], [1, 2, 7, 8]) == 2
    assert prefix_match([], [1, 2]) == 0 and prefix_match([1, 2], []) == 0
    assert prefix_match([5, IMG, IMG, 6, 7], [5, IMG, IMG, 6, 8], [IMG]) == 4   # same image, new question
    assert prefix_match([5, IMG, IMG, 6, 7], [5, IMG, IMG, 6, 8], [IMG], same_mm=False) == 0
    assert prefix_match([5, IMG, IMG, 6], [5, IMG, 6, 6], [IMG]) == 0           # diverges inside the image run
    assert prefix_match([5, 6, IMG, IMG], [5, 7, IMG, IMG], [IMG]) == 0         # image rows would sit at other positions
    assert prefix_match([5, 6, 7, IMG], [5, 6, 8], [IMG]) == 0                  # the cached tail still holds an image


@settings(max_examples=200, deadline=None)
@given(st.lists(st.integers(0, 5), max_size=24), st.lists(st.integers(0, 5), max_size=24), st.booleans())
def test_prefix_match_property(cached, ids, same_mm):
    from aha_b200.inference import prefix_match
    assert prefix_match(cached, ids, [4, 5], same_mm) == _prefix_rule(cached, ids, [4, 5], same_mm)


def test_mm_fingerprint_separates_requests():
    from aha_b200.inference import mm_fingerprint
    rng = np.random.default_rng(0)
    pv = rng.standard_normal((64, 1536)).astype(np.float32)
    grid = np.array([[1, 8, 8]], np.uint32)
    a = mm_fingerprint([pv, grid, None, None, None])
    assert a != 0 and a == mm_fingerprint([pv.copy(), grid.copy(), None, None, None])
    assert mm_fingerprint([None, None, None, None, None]) == 0 and mm_fingerprint(None) == 0
    pv2 = pv.copy(); pv2[63, 1535] = np.nextafter(pv2[63, 1535], np.float32(10))   # one ulp in the last element
    assert mm_fingerprint([pv2, grid, None, None, None]) != a
    assert mm_fingerprint([pv, np.array([[1, 4, 16]], np.uint32), None, None, None]) != a
    assert mm_fingerprint([None, None, pv, grid, None]) != a                        # the same bytes as a video are another request
    assert mm_fingerprint([pv.reshape(128, 768), grid, None, None, None]) != a
    big = rng.integers(0, 256, 3 * (1 << 20) + 12345, dtype=np.uint8)               # beyond 1 MiB: blocks hashed by several threads, folded in order
    fb = mm_fingerprint([big])
    assert fb == mm_fingerprint([big.copy()])
    for pos in (0, (1 << 20) - 1, 1 << 20, 2 * (1 << 20) + 77, big.size - 1):       # first / last byte of a block, the ragged tail
        y = big.copy(); y[pos] ^= 0x80
        assert mm_fingerprint([y]) != fb, pos
    sw = big.copy(); sw[:1 << 20], sw[1 << 20:2 << 20] = big[1 << 20:2 << 20].copy(), big[:1 << 20].copy()
    assert mm_fingerprint([sw]) != fb                                                # two blocks swapped: block order is part of the value
    for n in (0, 1, 7, 31, 32, 33, 100):                                             # every tail length of the 32-byte blocks
        x = np.arange(n, dtype=np.uint8)
        fp = mm_fingerprint([x])
        assert fp == mm_fingerprint([x.copy()])
        if n:
            y = x.copy(); y[-1] ^= 1
            assert mm_fingerprint([y]) != fp
