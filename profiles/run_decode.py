"""Small driver for ncu / timing experiments: VL2-shaped text stack, synthetic KV (no ViT), N decode steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aha_b200 import B200Model, synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
impl = int(os.environ.get("AHA_DECODE_IMPL", "0"))
ctx = int(os.environ.get("AHA_CTX", "2554"))
cfg = synth.get_config("qwen3", "q0.6")
shape = os.environ.get("AHA_SHAPE", "vl2")
if shape == "vl2":
    cfg.update(hidden_size=2048, intermediate_size=6144, rope_theta=5e6)   # Qwen3-VL-2B text stack shape (else: Qwen3-0.6B / ASR-0.6B)
w = synth.make_weights("qwen3", cfg, 0)
m = B200Model("qwen3", cfg, w, max_ctx=4096, max_prefill=64, decode_impl=impl)
del w
m.forward_initial(synth.synth_text_ids(8, 1000, 1), 0, want_logits=False)
# decode at offset ctx: KV pages below ctx hold whatever is in the pool (timing only)
toks, ms = m.decode_steps(5, ctx, steps, timed=True)
toks, ms = m.decode_steps(5, ctx, steps, timed=True)
print(f"shape={shape} impl={impl} pf={os.environ.get('AHA_FUSED_PF','-')} dbg={os.environ.get('AHA_FUSED_DBG','0')} ctx={ctx} steps={steps} ms/step={ms/steps:.4f} tok/s={1e3*steps/ms:.1f}")
if int(os.environ.get("AHA_FUSED_DBG", "0")) & 4:
    m.decode_steps(5, ctx, 1)
    c = m.debug_read("fused_trace", 0, 4096); p = m.debug_read("fused_trace", 1, 4096)
    print("consumer stamps", len(c), "producer stamps", len(p), "total us", c[-1] / 1e3, p[-1] / 1e3)
    # consumer: stamp0, then per layer: P1(load,gemv,bar) P2(start,attn,bar) P3(load,gemv,bar) P4(...) P5(...) => 15 per layer
    import numpy as np
    d = np.diff(c)[: 15 * 28].reshape(28, 15)
    names = ["P1 load_x", "P1 gemv", "P1 bar", "P2 -", "P2 attn", "P2 bar", "P3 load_x", "P3 gemv", "P3 bar", "P4 load_x", "P4 gemv", "P4 bar", "P5 load_x", "P5 gemv", "P5 bar"]
    for n, v in zip(names, d[2:].mean(0)):
        print(f"  consumer {n:10s} {v/1e3:7.2f} us")
    print("  consumer per layer", d[2:].sum(1).mean() / 1e3, "us; lm_head phase", (c[-1] - c[15 * 28]) / 1e3)
    dp = np.diff(p)[: 5 * 28].reshape(28, 5)
    for n, v in zip(["qkv", "attn", "o", "gate_up", "down"], dp[2:].mean(0)):
        print(f"  producer {n:8s} {v/1e3:7.2f} us")
    # one layer on a common time axis (us since kernel start)
    Lx = 10
    cb = c[1 + 15 * Lx: 1 + 15 * (Lx + 1)] / 1e3
    pb = p[1 + 5 * Lx: 1 + 5 * (Lx + 1)] / 1e3
    t0 = c[15 * Lx] / 1e3
    print(f"  layer {Lx} starts at {t0:.1f} us; consumer events (rel):", " ".join(f"{n.replace(' ', '')}@{v - t0:.1f}" for n, v in zip(names, cb)))
    print(f"  producer finished issuing (rel): ", " ".join(f"{n}@{v - t0:.1f}" for n, v in zip(["qkv", "attn", "o", "gate_up", "down"], pb)))

if int(os.environ.get("AHA_FUSED_DBG", "0")) & 64:
    m.decode_steps(5, ctx, 1)
    arr = np.stack([m.debug_read("fused_cta_trace", i, 256) for i in range(148)])
    os.makedirs("gpurun_out", exist_ok=True)
    np.save("gpurun_out/cta_trace.npy", arr)
    print("saved gpurun_out/cta_trace.npy", arr.shape)
