"""GEMM shape sweep through aha_b200_debug_gemm: the prefill shapes of the Qwen3-VL-2B workload, tcgen05 128x128 (impl 2) against
the persistent 128x256 kernel (impl 3).  Useful flops = 2*M*N*K (the kernel issues twice that: hi and lo activation halves)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aha_b200 import B200Model, synth
cfg = synth.get_config("qwen3", "tiny")
m = B200Model("qwen3", cfg, synth.make_weights("qwen3", cfg, 0), max_ctx=64)
SHAPES = [("vit qkv", 8160, 3072, 1024, 0), ("vit proj", 8160, 1024, 1024, 1), ("vit fc1", 8160, 4096, 1024, 2), ("vit fc2", 8160, 1024, 4096, 1),
          ("llm qkv", 2554, 4096, 2048, 0), ("llm o", 2554, 2048, 2048, 1), ("llm gate_up", 2554, 12288, 2048, 3), ("llm down", 2554, 2048, 6144, 1),
          ("8b gate_up", 16896, 24576, 4096, 3)]
rng = np.random.default_rng(0)
IMPLS = [int(v) for v in os.environ.get("AHA_GEMM_IMPLS", "2,3").split(",")]
for name, M, N, K, epi in SHAPES:
    x = rng.standard_normal((M, K)).astype(np.float32)
    w = (0.05 * rng.standard_normal((N, K))).astype(np.float16)
    bias = rng.standard_normal(N).astype(np.float32) if epi in (0, 2) else None
    resid = rng.standard_normal((M, N)).astype(np.float32) if epi == 1 else None
    line = f"{name:12s} M={M:5d} N={N:5d} K={K:4d} epi={epi}"
    ys = {}
    for impl in IMPLS:
        y, ms = m.debug_gemm(x, w, bias=bias, resid=resid, impl=impl, epi=epi, act=3 if epi == 2 else 0, iters=20)
        ys[impl] = y
        line += f" | impl {impl}: {ms / 20 * 1e3:8.1f} us {2.0 * M * N * K / (ms / 20 * 1e-3) / 1e12:6.1f} TF/s"
    line += f" | max|d| {max(np.abs(ys[IMPLS[0]] - ys[i]).max() for i in IMPLS):.2e}"
    print(line, flush=True)
