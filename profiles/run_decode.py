"""Driver for ncu / timing experiments on the decode step: text stack only (VL2 or Q0.6 shape), synthetic KV (no ViT).

  python profiles/run_decode.py [steps]              one configuration from the environment (AHA_DECODE_IMPL, AHA_FUSED_DBG, ...)
  python profiles/run_decode.py [steps] --sweep SPEC one process, many configurations (weights generated once):
      SPEC = ';'-separated runs, each ','-separated key=value with keys impl, dbg, stages, ctx, tl (1 = print the phase
      timeline), st (1 = dump the per-stage trace of an AHA_STAGE_TRACE build to gpurun_out/stage_trace_<tag>.npy)
      e.g. --sweep "impl=0;impl=0,dbg=3;impl=3,tl=1"
AHA_SHAPE = vl2 (default) | q0.6; AHA_LIB = path of an alternative libaha_b200.so (copied over the in-tree one by the caller)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aha_b200 import B200Model, synth

argv = [a for a in sys.argv[1:]]
sweep = None
if "--sweep" in argv:
    i = argv.index("--sweep")
    sweep = argv[i + 1]
    del argv[i:i + 2]
steps = int(argv[0]) if argv else 8
shape = os.environ.get("AHA_SHAPE", "vl2")
cfg = synth.get_config("qwen3", "q0.6")
if shape == "vl2":
    cfg.update(hidden_size=2048, intermediate_size=6144, rope_theta=5e6)   # Qwen3-VL-2B text stack shape (else: Qwen3-0.6B / ASR-0.6B)
w = synth.make_weights("qwen3", cfg, 0)
NAMES = ["P1 load_x", "P1 gemv", "P1 bar", "P2 -", "P2 attn", "P2 bar", "P3 load_x", "P3 gemv", "P3 bar", "P4 load_x", "P4 gemv", "P4 bar", "P5 load_x", "P5 gemv", "P5 bar"]


def timeline(m, ctx):
    m.decode_steps(5, ctx, 1)
    c = m.debug_read("fused_trace", 0, 4096); p = m.debug_read("fused_trace", 1, 4096)
    print("  consumer stamps", len(c), "producer stamps", len(p), "total us", c[-1] / 1e3, p[-1] / 1e3)
    d = np.diff(c)[: 15 * 28].reshape(28, 15)
    for n, v in zip(NAMES, d[2:].mean(0)):
        print(f"    consumer {n:10s} {v/1e3:7.2f} us")
    print("    consumer per layer", d[2:].sum(1).mean() / 1e3, "us; lm_head phase", (c[-1] - c[15 * 28]) / 1e3)
    dp = np.diff(p)[: 5 * 28].reshape(28, 5)
    for n, v in zip(["qkv", "attn", "o", "gate_up", "down"], dp[2:].mean(0)):
        print(f"    producer {n:8s} {v/1e3:7.2f} us")
    Lx = 10
    cb = c[1 + 15 * Lx: 1 + 15 * (Lx + 1)] / 1e3
    pb = p[1 + 5 * Lx: 1 + 5 * (Lx + 1)] / 1e3
    t0 = c[15 * Lx] / 1e3
    print(f"    layer {Lx} starts at {t0:.1f} us; consumer events (rel):", " ".join(f"{n.replace(' ', '')}@{v - t0:.1f}" for n, v in zip(NAMES, cb)))
    print(f"    producer finished issuing (rel): ", " ".join(f"{n}@{v - t0:.1f}" for n, v in zip(["qkv", "attn", "o", "gate_up", "down"], pb)))


def run(conf, models):
    impl = int(conf.get("impl", os.environ.get("AHA_DECODE_IMPL", "0")))
    ctx = int(conf.get("ctx", os.environ.get("AHA_CTX", "2554")))
    dbg = int(conf.get("dbg", os.environ.get("AHA_FUSED_DBG", "0")))
    if "stages" in conf:
        os.environ["AHA_FUSED_STAGES"] = str(conf["stages"])
    else:
        os.environ.pop("AHA_FUSED_STAGES", None)
    if impl not in models:
        m = B200Model("qwen3", cfg, w, max_ctx=4096, max_prefill=64, decode_impl=impl)
        m.forward_initial(synth.synth_text_ids(8, 1000, 1), 0, want_logits=False)
        models[impl] = m
    m = models[impl]
    os.environ["AHA_FUSED_DBG"] = str(dbg)
    # decode at offset ctx: KV pages below ctx hold whatever is in the pool (timing only)
    m.decode_steps(5, ctx, steps, timed=True)
    best = 1e9
    for _ in range(3):
        toks, ms = m.decode_steps(5, ctx, steps, timed=True)
        best = min(best, ms)
    print(f"shape={shape} impl={impl} dbg={dbg} stages={conf.get('stages', '-')} ctx={ctx} steps={steps} ms/step={best/steps:.4f} tok/s={1e3*steps/best:.1f}", flush=True)
    if int(conf.get("tl", 0)) or (sweep is None and dbg & 4):
        os.environ["AHA_FUSED_DBG"] = str(dbg | 4)
        timeline(m, ctx)
    if int(conf.get("st", 0)):
        os.makedirs("gpurun_out", exist_ok=True)
        for cta in (0, 77, 147):
            os.environ["AHA_FUSED_DBG"] = str(dbg | 128 | (cta << 16))
            m.debug_read("fused_stage_trace", 0, 4 * 8192)   # clears the region
            m.decode_steps(5, ctx, 1)
            t = m.debug_read("fused_stage_trace", 0, 4 * 8192).reshape(8192, 4)
            tag = f"{shape}_impl{impl}_dbg{dbg}_cta{cta}"
            np.save(f"gpurun_out/stage_trace_{tag}.npy", t)
            print(f"  saved gpurun_out/stage_trace_{tag}.npy ({int((t[:, 0] > 0).sum())} stages)")
    if int(conf.get("sy", 0)):   # sync anatomy (AHA_STAGE_TRACE build): barrier sub-steps and activation-load sub-steps of 3 CTAs
        os.makedirs("gpurun_out", exist_ok=True)
        for cta in (0, 77, 147):
            os.environ["AHA_FUSED_DBG"] = str(dbg | 1024 | (cta << 16))
            m.debug_read("fused_sync_trace", 0, 8192)
            m.decode_steps(5, ctx, 1)
            t = m.debug_read("fused_sync_trace", 0, 8192)
            tag = f"{shape}_impl{impl}_dbg{dbg}_cta{cta}"
            np.save(f"gpurun_out/sync_trace_{tag}.npy", t)
            b = t[:4096].reshape(512, 8)[:140]
            l = t[4096:].reshape(1024, 4)[:112]
            d = np.diff(b[8:, :6], axis=1)
            print(f"  cta {cta} barrier anatomy (us, mean over barriers 8..139): bar.sync {d[:,0].mean()/1e3:.2f} red.release {d[:,1].mean()/1e3:.2f} poll {d[:,2].mean()/1e3:.2f} fence {d[:,3].mean()/1e3:.2f} bar.sync {d[:,4].mean()/1e3:.2f}; polls/barrier {b[8:,6].mean():.1f}")
            dl = np.diff(l[8:], axis=1)
            print(f"  cta {cta} load_x anatomy: loads+sumsq {dl[:,0].mean()/1e3:.2f} bar.sync {dl[:,1].mean()/1e3:.2f} scale+sts+bar {dl[:,2].mean()/1e3:.2f}")
    if sweep is None and dbg & 64:
        m.decode_steps(5, ctx, 1)
        arr = np.stack([m.debug_read("fused_cta_trace", i, 256) for i in range(148)])
        os.makedirs("gpurun_out", exist_ok=True)
        np.save("gpurun_out/cta_trace.npy", arr)
        print("saved gpurun_out/cta_trace.npy", arr.shape)
    os.environ["AHA_FUSED_DBG"] = "0"


models = {}
if sweep is None:
    run({}, models)
else:
    for spec in sweep.split(";"):
        spec = spec.strip()
        if not spec:
            continue
        conf = dict(kv.split("=") for kv in spec.split(","))
        try:
            run(conf, models)
        except Exception as e:   # an experimental variant must not end the sweep
            print(f"shape={shape} {spec}: FAILED {type(e).__name__}: {e}", flush=True)
for m in models.values():
    m.close()
