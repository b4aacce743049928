#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: greedy decode tokens/s of Qwen3-VL-2B on a 1920x1080 image + 512-token
prompt, reported as absolute and as a fraction of the HBM roofline, with the reference's CPU path timed beside it.

A "step" is one greedy decode step (one token) of the hot path at ctx ~= 2554 + i, i.e. one replay of the
decode-step CUDA graph with the token fed back on the device.  The ViT + LLM prefill of the image prompt runs
before the timed region and is reported in `config`.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torch.distributed.run)
  python bench.py --impl reference ...                     (CPU port of the reference path on the host cores)

value  = K / (CUDA-event time of K graph replays, max over ranks)          -- inputs resident in HBM
e2e    = K / wall time of K aha_b200_forward_step calls (host token in, host argmax out per step)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "decode tokens/sec Qwen3-VL-2B 1080p+512ctx"
UNIT = "tokens/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload(preset):
    """Synthetic config-3 inputs: 1088x1920 noise image (resize = identity) + 512 text ids."""
    if preset == "vl2":
        return dict(kind="qwen3vl", preset="vl2", image=(1088, 1920), n_text=512, max_ctx=4096, max_patches=8192)
    if preset == "tiny":  # functional check of this script on small shapes (not a bench line)
        return dict(kind="qwen3vl", preset="tiny", image=(256, 320), n_text=64, max_ctx=1024, max_patches=1024)
    raise SystemExit(f"unknown preset {preset}")


def build_cpu_decoder(cfg, w, ctx, rope_delta):
    """The reference's CPU decode step (oracle port, numpy fp32 + BLAS threads): Qwen3-VL text stack with a
    synthetic KV cache of `ctx` tokens so that the ViT/prefill (minutes of CPU time) is not part of the sample."""
    from oracle.qwen3vl import Qwen3VLModel
    w32 = {k: (v.astype(np.float32) if k.startswith("model.language_model") else v) for k, v in w.items()}
    m = Qwen3VLModel(cfg, w32)
    tc = cfg["text_config"]
    rng = np.random.default_rng(0)
    for l in m.text.layers:
        shp = (1, tc["num_key_value_heads"], ctx, tc["head_dim"])
        l.attn.kv_cache = (rng.standard_normal(shp, dtype=np.float32), rng.standard_normal(shp, dtype=np.float32))
    m.rope_deltas = rope_delta
    return m


def cpu_decode_steps(m, ctx, n, tok=5):
    t0 = time.perf_counter()
    for i in range(n):
        logits = m.forward_step(np.array([[tok]]), ctx + i)
        tok = int(np.argmax(logits))
    return time.perf_counter() - t0


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of this run, written to the real stdout (fd 1 is pointed at stderr for everything else so that
    library chatter such as NCCL's version banner can never pollute it)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of one decode_step_fused_kernel launch, from the committed
    `ncu --set full` summary (profiles/r01_fused_decode_final_ncu_summary.csv)."""
    p = os.path.join(ROOT, "profiles", "r01_fused_decode_final_ncu_summary.csv")   # capture of the final tree of round 1
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "r01_fused_decode_ncu_summary.csv")
    try:
        tot = 0.0
        for row in open(p).read().splitlines()[1:]:
            name, unit, val = row.split(",")
            if name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(val) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
        return tot or None
    except Exception:
        return None


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--preset", default=os.environ.get("AHA_BENCH_PRESET", "vl2"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--decode-impl", type=int, default=int(os.environ.get("AHA_DECODE_IMPL", "0")))
    ap.add_argument("--parallelism", default=os.environ.get("AHA_PARALLELISM", "replicas"), choices=["replicas", "tp"],
                    help="N > 1: independent replicas (one request per GPU, default) or tensor parallelism of ONE request")
    args = ap.parse_args()
    K, W = args.steps, max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = workload(args.preset)

    from aha_b200 import synth
    cfg = synth.get_config(wl["kind"], wl["preset"])
    tc = cfg["text_config"]
    merge2 = cfg["vision_config"]["spatial_merge_size"] ** 2
    h, w_ = wl["image"]
    n_img_tok = (h // 16) * (w_ // 16) // merge2
    S = 1 + n_img_tok + 1 + wl["n_text"]
    config = {"workload": f"Qwen3-VL-2B shape ({wl['preset']}), random-init fp16 weights (seed 0), synthetic {w_}x{h} image "
                          f"({n_img_tok} image tokens) + {wl['n_text']} text ids, greedy decode at ctx {S}+",
              "prompt_tokens": S, "kv_dtype": "f32", "weight_dtype": "f16", "accumulate": "f32", "batch": 1,
              "l2_policy": "inputs larger than L2 (3.4 GB of weights streamed per step, 126 MB L2)",
              "parallelism": "single GPU" if world == 1 else (
                  f"tp{world}: one request, heads/MLP rows sharded, NCCL all-reduce after o_proj and down_proj" if args.parallelism == "tp"
                  else f"{world} independent replicas (one request per GPU, no data-path collective)")}

    # --------------------------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        log("[reference] generating weights ...")
        wts = synth.make_weights(wl["kind"], cfg, 0)
        grid = np.array([[1, h // 16, w_ // 16]])
        ids = synth.vl_prompt_ids(cfg, grid, wl["n_text"])
        from oracle.qwen3vl import get_rope_index
        _, delta = get_rope_index(ids, grid, cfg)
        m = build_cpu_decoder(cfg, wts, S, delta)
        cpu_decode_steps(m, S, W)
        dt = cpu_decode_steps(m, S + W, K)
        val = K / dt
        cores = os.cpu_count()
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"{K} greedy decode steps of the oracle port (numpy fp32, BLAS threads={cores}) of the "
                                           f"reference's Qwen3-VL text stack at ctx {S}+ with a synthetic KV cache; ViT+prefill not in the sample; "
                                           "the reference itself (Rust/Candle) cannot be built here (no cargo/rustc)"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    # --------------------------------------------------------------------------------- B200 arm
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(local_rank)

    from aha_b200 import B200Model
    t0 = time.time()
    wts = synth.make_weights(wl["kind"], cfg, 0)
    log(f"[rank {rank}] weights generated in {time.time() - t0:.1f}s")
    t0 = time.time()
    tp = world > 1 and args.parallelism == "tp"
    tp_kw = {}
    if tp:
        from aha_b200 import nccl_unique_id
        uid = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        tp_kw = dict(tp_rank=rank, tp_world=world, tp_unique_id=uid[0])
    m = B200Model(wl["kind"], cfg, wts, eos_ids=[], device=local_rank, max_ctx=wl["max_ctx"], max_prefill=wl["max_ctx"],
                  max_patches=wl["max_patches"], decode_impl=args.decode_impl, **tp_kw)
    log(f"[rank {rank}] model created in {time.time() - t0:.1f}s")
    img = synth.synth_image(h, w_, 1)
    pv, grid = m.image_patchify(img)
    ids = synth.vl_prompt_ids(cfg, grid, wl["n_text"])
    assert len(ids) == S
    data = [pv, grid, None, None, None]

    # request 1 through the public generate() (warms everything up, gives the prefill / ViT split)
    toks, usage = m.generate(ids, data, max_tokens=4)
    toks2, usage = m.generate(ids, data, max_tokens=4)
    assert toks == toks2
    log(f"[rank {rank}] prefill {usage['prompt_secs']:.3f}s (vision tower {usage['vision_secs']:.3f}s)")

    # prefill for the timed decode
    m.forward_initial(ids, 0, data, want_logits=False)
    tok = m.last_argmax
    rope_delta = int(m.debug_read("rope_delta", 0, 1)[0])
    warm = m.decode_steps(tok, S, W)
    m.reset_stats()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    # ---- device-resident value: K graph replays, token fed back on the device
    reps = max(1, int(os.environ.get("AHA_BENCH_REPS", "3")))
    best_ms = None
    barrier()
    wall0 = time.perf_counter()
    for _ in range(reps):
        out, ms = m.decode_steps(warm[-1], S + W, K, timed=True)
        best_ms = ms if best_ms is None else min(best_ms, ms)
    barrier()
    wall_value = time.perf_counter() - wall0
    st = m.stats()
    launches = st["kernel_launches"] // reps
    # ---- e2e: K forward_step calls through the C ABI, host token in / host argmax out each step
    t = out[-1] if out else tok
    barrier()
    e0 = time.perf_counter()
    for i in range(K):
        m.forward_step(np.array([t], np.uint32), S + W + i, want_logits=False)
        t = m.last_argmax
    barrier()
    e2e_s = time.perf_counter() - e0
    # ---- e2e with the full logits row returned to the host each step (what the reference's sampler consumes)
    e0 = time.perf_counter()
    for i in range(min(K, 32)):
        m.forward_step(np.array([t], np.uint32), S + W + i, want_logits=True)
    e2e_logits_s = (time.perf_counter() - e0) / min(K, 32)
    clocks = sampler.stop()

    # ---- dominant kernel roofline (gate/up GEMV: 41% of the step's bytes), timed alone with CUDA events
    peak, peak_src = measured_peaks()
    kernels = {}
    for name in ("gemv_gate_up", "gemv_down", "gemv_qkv", "gemv_o", "gemv_lm_head"):
        kms, kb = m.bench_kernel(name, 280 if name != "gemv_lm_head" else 20)
        kernels[name] = {"avg_us": kms * 1e3, "bytes": kb, "gbps": kb / (kms * 1e-3) / 1e9}
    fused = st["kernels_per_decode_step"] == 1

    from aha_b200 import dist_util
    dev = f"cuda:{local_rank}"
    jobs = 1 if tp else world                                                           # TP decodes ONE request on all ranks
    value, max_ms = dist_util.aggregate_throughput(K, best_ms, jobs, dist, dev)        # units of all ranks / slowest rank
    e2e_val, _ = dist_util.aggregate_throughput(K, e2e_s * 1e3, jobs, dist, dev)
    step_ms = max_ms / K
    avg_ctx = S + W + (K - 1) / 2.0 + 1
    step_bytes = st["decode_bytes_per_step_fixed"] + st["kv_bytes_per_token"] * (avg_ctx + 1)
    kv_read = st["kv_bytes_per_token"] * avg_ctx

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        log("[rank 0] timing the CPU port of the reference path ...")
        cm = build_cpu_decoder(cfg, wts, S, rope_delta)
        t1 = cpu_decode_steps(cm, S, 1)
        n = int(min(max(20.0 / max(t1, 1e-3), 2), 24))
        dt = cpu_decode_steps(cm, S + 1, n)
        cores = os.cpu_count()
        cpu_base = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{n} greedy decode steps of the oracle port (numpy fp32, BLAS threads={cores}) of the Qwen3-VL text stack "
                              f"at ctx {S}+ with a synthetic KV cache (ViT+prefill excluded); Rust/Candle reference not buildable here"}
        del cm

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": step_ms,
                "higher_is_better": True, "scaling": "strong" if tp else "weak", "vs_baseline": None,
                "dtype": "f32 (fp16 weights, fp32 activations/accumulate/KV)",
                "data": "synthetic", "config": dict(config, prefill_secs=usage["prompt_secs"], vision_secs=usage["vision_secs"],
                                                    kernels_per_step=st["kernels_per_decode_step"], reps=reps,
                                                    value_wall_check_s=wall_value),
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": 16, "d2h_bytes_per_step": 4,
                        "with_logits_d2h_tokens_per_s": 1.0 / e2e_logits_s, "logits_bytes": 4 * tc["vocab_size"]},
                "roofline": ({"bound": "hbm", "kernel": "decode_step_fused_kernel (the whole decode step: one launch per token)",
                              "achieved": step_bytes / (step_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": step_bytes / (step_ms * 1e-3) / 1e9 / peak, "peak_source": peak_src, "traffic": ncu_traffic_bytes(),
                              "bytes_per_launch": step_bytes, "avg_launch_us": step_ms * 1e3, "per_op_kernels": kernels}
                             if fused else
                             {"bound": "hbm", "kernel": "gemv_kernel<rmsnorm, swiglu> (gate/up projection)",
                              "achieved": kernels["gemv_gate_up"]["gbps"], "peak": peak, "unit": "GB/s",
                              "frac": kernels["gemv_gate_up"]["gbps"] / peak, "peak_source": peak_src, "traffic": None,
                              "bytes_per_launch": kernels["gemv_gate_up"]["bytes"], "avg_launch_us": kernels["gemv_gate_up"]["avg_us"],
                              "per_op_kernels": kernels}) | {
                             "step": {"bytes": step_bytes, "gbps": step_bytes / (step_ms * 1e-3) / 1e9,
                                      "frac_full": step_bytes / (step_ms * 1e-3) / 1e9 / peak,
                                      "roofline_full_tok_s": peak * 1e9 / step_bytes,
                                      "roofline_kv_tok_s": peak * 1e9 / kv_read}},
                "decode_impl": "fused persistent kernel" if fused else "per-op kernels (CUDA graph)",
                "cpu_baseline": cpu_base}
        emit(line)
    m.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
