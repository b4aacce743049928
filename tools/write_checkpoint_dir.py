#!/usr/bin/env python
"""Write a synthetic checkpoint DIRECTORY in the layout the reference's loaders read, so that the day a Rust box is available
the Candle goldens can be produced with the unmodified reference and the oracle / the CUDA path can finally be pinned:

    python tools/write_checkpoint_dir.py qwen3 tiny /tmp/ckpt_qwen3_tiny
    # on a box with cargo:   aha run -m qwen3-0.6b -w /tmp/ckpt_qwen3_tiny -i "t17 t4 t250 t9"      (temperature 0 in generation_config.json)
    # tests/golden/make_golden.py holds the same weights (aha_b200.synth, seed 0) and the same ids.

Files (what `Qwen3GenerateModel::init` opens, /root/reference/src/models/qwen3/generate.rs:22-50; VL: qwen3vl/generate.rs:33-63 +
qwen3vl/processor.rs:70-84; ASR: qwen3_asr/generate.rs:51-87):
  config.json               every field of Qwen3Config / Qwen3VLConfig / Qwen3AsrConfig is REQUIRED by serde (qwen3/config.rs:5-28)
  generation_config.json    Qwen3GenerationConfig (qwen3/config.rs:30-44): temperature 0 -> Sampling::ArgMax
  model.safetensors         fp16 tensors under the checkpoint names (aha_b200.synth), "torch_dtype": "float16"
  tokenizer.json            a WordLevel tokenizer whose token "t<i>" has id i (whitespace pre-tokenizer): prompts spell ids
  tokenizer_config.json     chat_template that emits the message contents verbatim (ChatTemplate::init reads it from here)
  preprocessor_config.json / video_preprocessor_config.json   (VL / ASR processors)
The reference defaults CPU runs to F16 for float16 checkpoints (utils/mod.rs:102-112); pass dtype F32 for the oracle's arithmetic."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aha_b200 import synth  # noqa: E402


def _text_extras(tc):
    return dict(attention_dropout=0.0, bos_token_id=tc.get("bos_token_id", 0), initializer_range=0.02, max_position_embeddings=40960,
                max_window_layers=tc["num_hidden_layers"], torch_dtype="float16", use_cache=True, use_sliding_window=False)


def write_tokenizer(out, vocab_size, specials=()):
    from tokenizers import Tokenizer, models, pre_tokenizers
    vocab = {f"t{i}": i for i in range(vocab_size)}
    for name, i in specials:
        vocab.pop(f"t{i}", None)
        vocab[name] = i
    tok = Tokenizer(models.WordLevel(vocab=vocab, unk_token="t0"))
    tok.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    tok.save(os.path.join(out, "tokenizer.json"))
    json.dump({"chat_template": "{%- for m in messages -%}{{ m.content }}{%- endfor -%}", "model_max_length": 131072},
              open(os.path.join(out, "tokenizer_config.json"), "w"), indent=1)


def main():
    if len(sys.argv) != 4:
        raise SystemExit(__doc__)
    kind, preset, out = sys.argv[1:]
    from safetensors.numpy import save_file
    os.makedirs(out, exist_ok=True)
    cfg = synth.get_config(kind, preset)
    w = synth.make_weights(kind, cfg, 0)
    save_file({k: np.ascontiguousarray(v) for k, v in w.items()}, os.path.join(out, "model.safetensors"))
    specials = []
    if kind == "qwen3":
        full = dict(cfg, **_text_extras(cfg))
        eos = [cfg["eos_token_id"]]
        vocab = cfg["vocab_size"]
    elif kind == "qwen3vl":
        tc = dict(cfg["text_config"], **_text_extras(cfg["text_config"]))
        vc = dict(cfg["vision_config"], initializer_range=0.02, model_type="qwen3_vl", depth=cfg["vision_config"]["depth"])
        full = dict(cfg, text_config=tc, vision_config=vc, architectures=["Qwen3VLForConditionalGeneration"], model_type="qwen3_vl")
        eos = [cfg["text_config"]["eos_token_id"]]
        vocab = cfg["text_config"]["vocab_size"]
        specials = [("<|image_pad|>", cfg["image_token_id"]), ("<|video_pad|>", cfg["video_token_id"]),
                    ("<|vision_start|>", cfg["vision_start_token_id"]), ("<|vision_end|>", cfg["vision_end_token_id"])]
        pre = dict(size=dict(shortest_edge=65536, longest_edge=16777216), patch_size=16, temporal_patch_size=2, merge_size=2,
                   image_mean=[0.5, 0.5, 0.5], image_std=[0.5, 0.5, 0.5], do_resize=True, do_rescale=True, do_normalize=True)
        json.dump(pre, open(os.path.join(out, "preprocessor_config.json"), "w"), indent=1)
        json.dump(dict(pre, size=dict(shortest_edge=4096, longest_edge=25165824), fps=2, min_frames=4, max_frames=768),
                  open(os.path.join(out, "video_preprocessor_config.json"), "w"), indent=1)
    else:
        tk = cfg["thinker_config"]
        tc = dict(tk["text_config"], **_text_extras(tk["text_config"]))
        full = dict(cfg, thinker_config=dict(tk, text_config=tc))
        eos = [tk["text_config"]["eos_token_id"]]
        vocab = tk["text_config"]["vocab_size"]
        specials = [("<|audio_pad|>", tk["audio_token_id"]), ("<|audio_start|>", tk["audio_start_token_id"]), ("<|audio_end|>", tk["audio_end_token_id"])]
        json.dump(dict(feature_size=128, hop_length=160, n_fft=400, sampling_rate=16000, chunk_length=30, padding_value=0.0),
                  open(os.path.join(out, "preprocessor_config.json"), "w"), indent=1)
    json.dump(full, open(os.path.join(out, "config.json"), "w"), indent=1)
    json.dump(dict(bos_token_id=0, pad_token_id=0, do_sample=False, eos_token_id=eos, top_p=1.0, top_k=1, temperature=0.0, repetition_penalty=1.0),
              open(os.path.join(out, "generation_config.json"), "w"), indent=1)
    write_tokenizer(out, vocab, specials)
    print(f"wrote {kind}/{preset} to {out}: {sorted(os.listdir(out))}")


if __name__ == "__main__":
    main()
