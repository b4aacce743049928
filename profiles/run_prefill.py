"""Prefill-only driver for ncu launch lists: Qwen3-VL-2B shape, 1088x1920 image + 512 text ids, one forward_initial."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aha_b200 import B200Model, synth
cfg = synth.get_config("qwen3vl", "vl2")
w = synth.make_weights("qwen3vl", cfg, 0)
m = B200Model("qwen3vl", cfg, w, max_ctx=4096, max_prefill=4096, max_patches=8192, attn_impl=int(os.environ.get("AHA_ATTN_IMPL", "0")),
              gemm_impl=int(os.environ.get("AHA_GEMM_IMPL", "0")))
del w
pv, grid = m.image_patchify(synth.synth_image(1088, 1920, 1))
ids = synth.vl_prompt_ids(cfg, grid, 512)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    m.clear_cache()
    t0 = time.perf_counter()
    m.forward_initial(ids, 0, [pv, grid, None, None, None], want_logits=False)
    print(f"prefill {time.perf_counter() - t0:.4f} s")
