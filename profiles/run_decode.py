"""Small driver for ncu / timing experiments: VL2-shaped text stack, synthetic KV (no ViT), N decode steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aha_b200 import B200Model, synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
impl = int(os.environ.get("AHA_DECODE_IMPL", "0"))
ctx = int(os.environ.get("AHA_CTX", "2554"))
cfg = synth.get_config("qwen3", "q0.6")
cfg.update(hidden_size=2048, intermediate_size=6144, rope_theta=5e6)   # Qwen3-VL-2B text stack shape
w = synth.make_weights("qwen3", cfg, 0)
m = B200Model("qwen3", cfg, w, max_ctx=4096, max_prefill=64, decode_impl=impl)
del w
m.forward_initial(synth.synth_text_ids(8, 1000, 1), 0, want_logits=False)
# decode at offset ctx: KV pages below ctx hold whatever is in the pool (timing only)
toks, ms = m.decode_steps(5, ctx, steps, timed=True)
toks, ms = m.decode_steps(5, ctx, steps, timed=True)
print(f"impl={impl} dbg={os.environ.get('AHA_FUSED_DBG','0')} ctx={ctx} steps={steps} ms/step={ms/steps:.4f} tok/s={1e3*steps/ms:.1f}")
