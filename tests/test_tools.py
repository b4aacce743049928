"""tools/write_checkpoint_dir.py: the synthetic checkpoint directory carries every file and every serde-required field the
reference's loader opens (/root/reference/src/models/qwen3/generate.rs:22-50, qwen3/config.rs:5-44), so Candle goldens can be
produced with the unmodified reference on a box that has a Rust toolchain."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT

QWEN3_CONFIG_FIELDS = ["attention_bias", "attention_dropout", "bos_token_id", "eos_token_id", "head_dim", "hidden_act", "hidden_size",
                       "initializer_range", "intermediate_size", "max_position_embeddings", "max_window_layers", "num_attention_heads",
                       "num_hidden_layers", "num_key_value_heads", "rms_norm_eps", "rope_theta", "tie_word_embeddings", "torch_dtype",
                       "use_cache", "use_sliding_window", "vocab_size"]
GENERATION_FIELDS = ["bos_token_id", "pad_token_id", "do_sample", "eos_token_id", "top_p", "top_k", "temperature"]


def test_qwen3_checkpoint_dir(tmp_path):
    out = str(tmp_path / "ckpt")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "write_checkpoint_dir.py"), "qwen3", "tiny", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cfg = json.load(open(os.path.join(out, "config.json")))
    assert all(k in cfg for k in QWEN3_CONFIG_FIELDS)
    gen = json.load(open(os.path.join(out, "generation_config.json")))
    assert all(k in gen for k in GENERATION_FIELDS) and gen["temperature"] == 0.0 and isinstance(gen["eos_token_id"], list)
    from safetensors.numpy import load_file
    from aha_b200 import synth
    w = load_file(os.path.join(out, "model.safetensors"))
    want = synth.make_weights("qwen3", synth.get_config("qwen3", "tiny"), 0)
    assert set(w) == set(want) and all(np.array_equal(w[k], want[k]) for k in want)
    from tokenizers import Tokenizer
    assert Tokenizer.from_file(os.path.join(out, "tokenizer.json")).encode("t17 t4 t250 t9").ids == [17, 4, 250, 9]
    assert "chat_template" in json.load(open(os.path.join(out, "tokenizer_config.json")))
