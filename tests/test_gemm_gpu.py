"""Unit tests of the Linear-layer GEMM kernels through aha_b200_debug_gemm: the tcgen05 split-fp16 kernel (impl=2)
and the SIMT fp32 kernel (impl=1) against a float64 numpy product of the same fp16 weights."""
import math

import numpy as np
import pytest

from conftest import make_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m():
    cfg, w, model = make_model("qwen3", "tiny", max_ctx=64)
    yield model
    model.close()


def _ref(x, w16, bias, resid, epi, act):
    y = x.astype(np.float64) @ w16.astype(np.float64).T
    if bias is not None:
        y = y + bias
    if epi == 1:
        y = y + resid
    if epi == 2:
        if act == 1:
            y = y / (1 + np.exp(-y))
        elif act == 2:
            y = 0.5 * y * (1 + np.vectorize(math.erf)(y / math.sqrt(2)))
        elif act == 3:
            y = 0.5 * y * (1 + np.tanh(math.sqrt(2 / math.pi) * (y + 0.044715 * y ** 3)))
    if epi == 3:
        g, u = y[:, 0::2], y[:, 1::2]
        y = g / (1 + np.exp(-g)) * u
    return y


@pytest.mark.parametrize("impl", [1, 2, 3])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 256, 128), (1, 128, 64), (333, 384, 512), (129, 160, 1024), (64, 96, 2048)])
def test_gemm_store(m, impl, M, N, K):
    rng = np.random.default_rng(M * 7 + N + K)
    x = rng.standard_normal((M, K)).astype(np.float32)
    w = (0.05 * rng.standard_normal((N, K))).astype(np.float16)
    y, _ = m.debug_gemm(x, w, impl=impl)
    ref = _ref(x, w, None, None, 0, 0)
    assert np.abs(y - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("impl", [1, 2, 3])
@pytest.mark.parametrize("epi,act", [(1, 0), (2, 1), (2, 2), (2, 3), (3, 0)])
def test_gemm_epilogues(m, impl, epi, act):
    rng = np.random.default_rng(epi * 10 + act)
    M, N, K = 150, 256, 192
    x = rng.standard_normal((M, K)).astype(np.float32)
    w = (0.05 * rng.standard_normal((N, K))).astype(np.float16)
    bias = rng.standard_normal(N).astype(np.float32) if epi != 3 else None
    resid = rng.standard_normal((M, N)).astype(np.float32) if epi == 1 else None
    y, _ = m.debug_gemm(x, w, bias=bias, resid=resid, impl=impl, epi=epi, act=act)
    ref = _ref(x, w, bias, resid, epi, act)
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= 3e-5 * max(1.0, np.abs(ref).max())


def test_tcgen05_keeps_fp32_grade_accuracy_on_wide_range_inputs(m):
    """The hi/lo split must not lose the low bits of activations: mix magnitudes from 1e-3 to 1e3."""
    rng = np.random.default_rng(3)
    M, N, K = 256, 128, 1024
    x = (rng.standard_normal((M, K)) * np.exp(rng.uniform(-7, 7, (M, K)))).astype(np.float32)
    w = (0.05 * rng.standard_normal((N, K))).astype(np.float16)
    y, _ = m.debug_gemm(x, w, impl=2)
    ref = x.astype(np.float64) @ w.astype(np.float64).T
    scale = np.sqrt((x.astype(np.float64) ** 2) @ (w.astype(np.float64) ** 2).T)   # per-output error scale
    err = (np.abs(y - ref) / scale).max()
    # the split removes the operand rounding; what remains is the tensor core's fp32 accumulation (~1e-5 at K = 1024)
    assert err < 3e-5
    y16 = (x.astype(np.float16).astype(np.float64)) @ w.astype(np.float64).T          # what plain fp16 activations would give
    assert (np.abs(y16 - ref) / scale).max() > 8 * err


def test_gemm_tc_timing_report(m):
    rng = np.random.default_rng(0)
    M, N, K = 2560, 4096, 2048
    x = rng.standard_normal((M, K)).astype(np.float32)
    w = (0.02 * rng.standard_normal((N, K))).astype(np.float16)
    y2, ms2 = m.debug_gemm(x, w, impl=2, iters=10)
    y1, ms1 = m.debug_gemm(x, w, impl=1, iters=3)
    assert np.abs(y1 - y2).max() <= 2e-5 * np.abs(y1).max()
    fl = 2.0 * M * N * K
    print(f"\nGEMM {M}x{N}x{K}: tcgen05(split, incl. split pass) {ms2 / 10:.3f} ms = {fl / (ms2 / 10) / 1e9:.0f} TFLOP/s useful; "
          f"SIMT {ms1 / 3:.3f} ms = {fl / (ms1 / 3) / 1e9:.0f} TFLOP/s")
