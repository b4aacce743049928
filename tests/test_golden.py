"""Committed golden fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py from the oracle):
 - CPU: the oracle still reproduces them (guards the oracle against drift);
 - GPU: the CUDA path reproduces them through the C ABI."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, TOL, make_model, make_oracle
from aha_b200 import synth


def _load(name):
    return np.load(os.path.join(GOLDEN, name))


def test_oracle_reproduces_qwen3_golden():
    g = _load("qwen3_tiny.npz")
    cfg = synth.get_config("qwen3", "tiny")
    o = make_oracle("qwen3", cfg, synth.make_weights("qwen3", cfg, 0))
    got = o.forward_initial(g["ids"].reshape(1, -1), 0)[0, 0]
    assert np.abs(got - g["prefill_logits"]).max() < 1e-5
    for i, t in enumerate(g["forced"]):
        assert np.abs(o.forward_step(np.array([[t]]), 45 + i)[0, 0] - g["step_logits"][i]).max() < 1e-5


def test_oracle_reproduces_vl_and_asr_golden():
    g = _load("qwen3vl_tiny.npz")
    cfg = synth.get_config("qwen3vl", "tiny")
    from oracle.qwen3vl import process_image
    o = make_oracle("qwen3vl", cfg, synth.make_weights("qwen3vl", cfg, 0))
    pv, grid = process_image(g["image"])
    assert abs(float(pv.astype(np.float64).sum()) - float(g["pixel_checksum"])) < 1e-3
    got = o.forward_initial(g["ids"].reshape(1, -1), 0, [pv, grid, None, None, None])[0, 0]
    assert np.abs(got - g["prefill_logits"]).max() < 1e-5 and o.rope_deltas == int(g["rope_delta"])
    a = _load("qwen3_asr_tiny.npz")
    cfg = synth.get_config("qwen3_asr", "tiny")
    o = make_oracle("qwen3_asr", cfg, synth.make_weights("qwen3_asr", cfg, 0))
    from oracle.audio import WhisperFeatureExtractor
    mel = WhisperFeatureExtractor().call(a["wave"][None], 16000)[0]
    assert np.abs(mel - a["mel"]).max() < 1e-5
    assert np.abs(o.forward_initial(a["ids"].reshape(1, -1), 0, [mel])[0, 0] - a["prefill_logits"]).max() < 1e-5


@pytest.mark.gpu
def test_cuda_reproduces_qwen3_golden():
    g = _load("qwen3_tiny.npz")
    cfg, w, m = make_model("qwen3", "tiny", max_ctx=256)
    try:
        got = m.forward_initial(g["ids"], 0)[0, 0]
        assert np.abs(got - g["prefill_logits"]).max() <= TOL
        for i, t in enumerate(g["forced"]):
            assert np.abs(m.forward_step(np.array([t], np.uint32), 45 + i)[0, 0] - g["step_logits"][i]).max() <= TOL
        m.clear_cache()
        toks, _ = m.generate(g["ids"], max_tokens=24)
        assert toks == g["greedy"].tolist()
    finally:
        m.close()


@pytest.mark.gpu
def test_cuda_reproduces_vl_and_asr_golden():
    g = _load("qwen3vl_tiny.npz")
    cfg, w, m = make_model("qwen3vl", "tiny", max_ctx=512, max_patches=1024)
    try:
        pv, grid = m.image_patchify(g["image"])                      # GPU patchify feeds the GPU model
        assert grid.tolist() == g["grid"].tolist()
        got = m.forward_initial(g["ids"], 0, [pv, grid, None, None, None])[0, 0]
        assert np.abs(got - g["prefill_logits"]).max() <= TOL
        assert int(m.debug_read("rope_delta", 0, 1)[0]) == int(g["rope_delta"])
        assert np.abs(m.forward_step(np.array([5], np.uint32), len(g["ids"]))[0, 0] - g["step_logits"]).max() <= TOL
    finally:
        m.close()
    a = _load("qwen3_asr_tiny.npz")
    cfg, w, m = make_model("qwen3_asr", "tiny", max_ctx=512, max_frames=600)
    try:
        mel = m.mel_spectrogram(a["wave"])                            # GPU mel feeds the GPU model
        assert np.abs(mel - a["mel"]).max() <= 1e-3
        got = m.forward_initial(a["ids"], 0, [mel])[0, 0]
        assert np.abs(got - a["prefill_logits"]).max() <= TOL
    finally:
        m.close()
