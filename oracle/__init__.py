"""CPU oracle: a numpy fp32 restatement of the jhqxxx/aha Qwen3 / Qwen3-VL / Qwen3-ASR
prefill+decode path (reference @ e29ddc5).

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it.  The product path (`aha_b200/`, `libaha_b200.so`) never does, and has
no CPU fallback.

Parity status: **parity unpinned by the reference** -- aha's own tests contain no
assertions and no golden vectors, and the arithmetic lives in the
un-vendored crates candle-core / candle-nn / candle-transformers 0.9.2 (Cargo.toml:10-12),
gemm 0.18, realfft 3.5.  The restatement is therefore anchored on
  (1) the reference's call sites, line by line (each function cites file:line),
  (2) closed-form known answers derivable from the source (tests/test_oracle_kat.py),
  (3) cross-checks against HF transformers 5.5 (tests/test_oracle_hf.py): Qwen3, Qwen3-VL, the
      image patchify, the ASR audio tower and the log-mel frontend, the reference's deliberate
      deviations from HF switched off for the comparison and covered by known-answer tests.
Candle op semantics assumed (published behaviour of candle 0.9.x):
  Linear = x @ W^T (+ b); RmsNorm = x / sqrt(mean(x^2) + eps) * w with f32 statistics;
  LayerNorm = (x - mean) / sqrt(var + eps) * w + b; softmax_last_dim is max-subtracted;
  Activation::Silu = x * sigmoid(x); Activation::Gelu = erf GELU;
  Activation::GeluPytorchTanh / Tensor::gelu() = tanh-approximated GELU;
  Embedding = row gather; Sampling::ArgMax = first maximal index;
  f32 -> u32 to_dtype truncates toward zero.
"""
