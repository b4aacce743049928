"""Opt-in variants that have not been validated on hardware yet.  They are NOT on any default path; these tests only run
when AHA_TEST_EXPERIMENTAL=1 so that an unfinished experiment can never turn the GPU suite red.

GQA groups 1 and 4: instantiations of the decode attention (fused and per-op) that no shipped model shape reaches.

decode_impl = 3: the fused decode kernel with the K-split down projection (decode_fused.cuh, variant KS): gate/up and
down in one phase, fp32 reductions into a global accumulator, 4 grid barriers per layer instead of 5.
decode_impl = 4: variant KO, K-split o_proj behind a kv-group barrier (each CTA merges only its own group's partials).
decode_impl = 5: both."""
import os

import numpy as np
import pytest

from conftest import TOL, make_model, make_oracle

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("AHA_TEST_EXPERIMENTAL") != "1", reason="set AHA_TEST_EXPERIMENTAL=1 to run unvalidated variants")]


def _ids(n, vocab, seed):
    return np.random.default_rng(seed).integers(0, min(vocab, 1000), n).astype(np.uint32)


@pytest.mark.parametrize("impl", [3, 4, 5])
@pytest.mark.parametrize("preset", ["tiny", "mid"])
def test_ksplit_decode_matches_the_default_fused_kernel_and_the_oracle(preset, impl):
    cfg, w, m = make_model("qwen3", preset, max_ctx=512)
    _, _, k = make_model("qwen3", preset, max_ctx=512, decode_impl=impl)
    o = make_oracle("qwen3", cfg, w)
    try:
        ids = _ids(50, cfg["vocab_size"], 8)
        for x in (m, k):
            x.clear_cache()
            x.forward_initial(ids, 0)
        o.clear_cache()
        o.forward_initial(ids.reshape(1, -1), 0)
        tok = 5
        for step in range(12):                       # teacher-forced: the same token feeds all three
            lm = m.forward_step(np.array([tok], np.uint32), 50 + step)[0, 0]
            lk = k.forward_step(np.array([tok], np.uint32), 50 + step)[0, 0]
            lo = o.forward_step(np.array([[tok]]), 50 + step)[0, 0]
            assert np.abs(lk - lo).max() <= TOL, (step, np.abs(lk - lo).max())
            assert np.abs(lk - lm).max() <= 1e-4, (step, np.abs(lk - lm).max())
            tok = int(np.argmax(lo))
        assert k.stats()["kernels_per_decode_step"] == 1
        a = m.decode_steps(5, 62, 24)
        b = k.decode_steps(5, 62, 24)
        assert list(a) == list(b)                    # greedy ids agree (ties aside, fp32 reduction order differs)
    finally:
        m.close(); k.close()


@pytest.mark.parametrize("preset", ["tiny-g1", "tiny-g4"])
@pytest.mark.parametrize("impl", [1, 2])
def test_decode_with_other_gqa_groups(preset, impl):
    cfg, w, m = make_model("qwen3", preset, max_ctx=512, decode_impl=impl)
    o = make_oracle("qwen3", cfg, w)
    try:
        ids = _ids(70, cfg["vocab_size"], 11)
        got = m.forward_initial(ids, 0)[0, 0]
        want = o.forward_initial(ids.reshape(1, -1), 0)[0, 0]
        assert np.abs(got - want).max() <= TOL
        tok = int(np.argmax(want))
        for step in range(16):
            lg = m.forward_step(np.array([tok], np.uint32), 70 + step)[0, 0]
            lo = o.forward_step(np.array([[tok]]), 70 + step)[0, 0]
            assert np.abs(lg - lo).max() <= TOL, (step, np.abs(lg - lo).max())
            tok = int(np.argmax(lo))
    finally:
        m.close()
