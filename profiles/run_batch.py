"""Lockstep batch decode (aha_b200_generate_batch) on the Qwen3-VL-2B text stack shape: aggregate decode tokens/s for 1..8 requests, beside
the same requests served one by one.  python profiles/run_batch.py [n_prompt] [n_gen]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aha_b200 import B200Model, synth

n_prompt = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n_gen = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg = synth.get_config("qwen3", "q0.6")
if os.environ.get("AHA_SHAPE", "vl2") == "vl2":
    cfg.update(hidden_size=2048, intermediate_size=6144, rope_theta=5e6)
w = synth.make_weights("qwen3", cfg, 0)
m = B200Model("qwen3", cfg, w, eos_ids=[], max_ctx=8 * (n_prompt + n_gen + 64), max_prefill=max(n_prompt + 32, 64))
del w
reqs = [dict(input_ids=synth.synth_text_ids(n_prompt + 3 * i, 151000, 500 + i), max_tokens=n_gen) for i in range(8)]
singles = []
t_single = 0.0
for r in reqs:
    t, u = m.generate(r["input_ids"], max_tokens=n_gen)
    singles.append(t); t_single += u["completion_secs"]
print(f"one by one (fused step kernel): {sum(len(t) - 1 for t in singles) / t_single:.1f} tok/s")
for impl in ("0", "2"):
    os.environ["AHA_BATCH_GEMV"] = impl
    for nb in (1, 2, 4, 8):
        m.generate_batch(reqs[:nb])
        res = m.generate_batch(reqs[:nb])
        dec = sum(len(t) - 1 for t, _ in res)
        secs = max(u["completion_secs"] for _, u in res)
        same = all(a == b[0] for a, b in zip(singles, res))
        print(f"batch {nb} ({'GEMV v1 (weights in registers)' if impl == '0' else 'GEMV v2 (cp.async weight ring)'}): {dec / secs:8.1f} tok/s aggregate, {1e3 * secs / (n_gen - 1):.3f} ms/step, ids equal to single runs: {same}")
m.close()
