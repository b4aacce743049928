"""Generates the committed golden fixtures from the oracle (run in the build container:
`python tests/golden/make_golden.py`).  Weights are NOT stored: they are re-derived from the seed by
aha_b200.synth; the fixtures hold inputs + oracle outputs (logits, greedy ids, gaps) for the tiny configs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from aha_b200 import synth  # noqa: E402
from oracle.audio import WhisperFeatureExtractor, get_feat_extract_output_lengths  # noqa: E402
from oracle.generate import GenerationContext, generate_generic  # noqa: E402
from oracle.qwen3 import Qwen3Model  # noqa: E402
from oracle.qwen3_asr import Qwen3ASRModel  # noqa: E402
from oracle.qwen3vl import Qwen3VLModel, process_image  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def gap(l):
    s = np.sort(np.asarray(l).reshape(-1))
    return float(s[-1] - s[-2])


def qwen3():
    cfg = synth.get_config("qwen3", "tiny")
    w = synth.make_weights("qwen3", cfg, 0)
    m = Qwen3Model(cfg, w, [cfg["eos_token_id"]])
    ids = synth.synth_text_ids(45, cfg["vocab_size"] - 8, 17)
    prefill = m.forward_initial(ids.reshape(1, -1), 0)[0, 0]
    forced = synth.synth_text_ids(6, cfg["vocab_size"] - 8, 18)
    steps = np.stack([m.forward_step(forced[i:i + 1].reshape(1, 1), 45 + i)[0, 0] for i in range(6)])
    m.clear_cache()
    ctx = GenerationContext(temperature=0.0, initial_seq_len=45, max_tokens=24)
    gen, _, _ = generate_generic(m, ids.reshape(1, -1), None, ctx)
    np.savez_compressed(os.path.join(OUT, "qwen3_tiny.npz"), ids=ids, prefill_logits=prefill, forced=forced, step_logits=steps,
                        greedy=np.array(gen, np.uint32), prefill_gap=gap(prefill))


def qwen3vl():
    cfg = synth.get_config("qwen3vl", "tiny")
    w = synth.make_weights("qwen3vl", cfg, 0)
    m = Qwen3VLModel(cfg, w, [cfg["text_config"]["eos_token_id"]])
    img = synth.synth_image(256, 320, 1)
    pv, grid = process_image(img)
    ids = synth.vl_prompt_ids(cfg, grid, 12)
    logits = m.forward_initial(ids.reshape(1, -1), 0, [pv, grid, None, None, None])[0, 0]
    delta = m.rope_deltas
    step = m.forward_step(np.array([[5]]), len(ids))[0, 0]
    np.savez_compressed(os.path.join(OUT, "qwen3vl_tiny.npz"), image=img, grid=grid, ids=ids, prefill_logits=logits, rope_delta=delta,
                        step_logits=step, pixel_checksum=np.float64(pv.astype(np.float64).sum()))


def qwen3_asr():
    cfg = synth.get_config("qwen3_asr", "tiny")
    w = synth.make_weights("qwen3_asr", cfg, 0)
    m = Qwen3ASRModel(cfg, w)
    wave = synth.synth_audio(2.5)
    mel = WhisperFeatureExtractor().call(wave[None], 16000)[0]
    ids = synth.asr_prompt_ids(cfg, get_feat_extract_output_lengths(mel.shape[1]))
    logits = m.forward_initial(ids.reshape(1, -1), 0, [mel])[0, 0]
    np.savez_compressed(os.path.join(OUT, "qwen3_asr_tiny.npz"), wave=wave.astype(np.float32), mel=mel.astype(np.float32), ids=ids,
                        prefill_logits=logits)


if __name__ == "__main__":
    qwen3(); qwen3vl(); qwen3_asr()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")
