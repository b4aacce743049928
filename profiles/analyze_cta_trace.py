"""Skew analysis of the fused decode kernel's grid barriers.

Input: gpurun_out/cta_trace.npy written by `AHA_FUSED_DBG=64 python profiles/run_decode.py 16` -- [148 CTAs][256]: the
%globaltimer stamp (ns, relative to CTA 0's first) of every CTA's ARRIVAL at every grid barrier, [:, 255] = %smid.
Prints, per barrier kind (P1..P5 of a layer), the spread of the arrivals, which CTAs arrive last, and the phase durations.
"""
import collections
import sys

import numpy as np

a = np.load(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/cta_trace.npy")
per_layer = int(sys.argv[2]) if len(sys.argv) > 2 else 5          # grid barriers per layer (5 default kernel, 4 variant KS)
smid = a[:, 255].astype(int)
n = ((a[0, :255] > -1e29).sum() // per_layer) * per_layer           # whole layers only
t = a[:, :n] / 1e3                                                   # us
last, first, med = t.max(0), t.min(0), np.median(t, 0)
names = [f"P{k + 1}" for k in range(per_layer)]
print(f"{a.shape[0]} CTAs on {len(set(smid))} SMs, {n} barriers")
for k, name in enumerate(names):
    sp = (last - first)[k::per_layer]
    lag = (last - med)[k::per_layer]
    late = (t - med[None])[:, k::per_layer]
    who = collections.Counter(t.argmax(0)[k::per_layer].tolist()).most_common(4)
    print(f"{name}: spread last-first {sp.mean():5.2f} us, last vs median {lag.mean():5.2f} us, lateness pct50/90/99/max "
          f"{np.percentile(late, 50):.2f}/{np.percentile(late, 90):.2f}/{np.percentile(late, 99):.2f}/{late.max():.2f}, last arrivers {who}")
d = np.diff(last)
print("phase durations (us, last arrival to last arrival):", [round(float(d[(k - 1) % per_layer::per_layer].mean()), 2) for k in range(per_layer)])
