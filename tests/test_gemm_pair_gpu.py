"""The CTA-pair tcgen05 GEMM (gemm_impl=4: cta_group::2, one 256 x 256 tile per cluster of two CTAs) against a float64 numpy product of the
same fp16 weights, and bit-for-bit against the single-CTA tensor-core kernel (same products, same accumulation order per output)."""
import numpy as np
import pytest

from conftest import make_model
from test_gemm_gpu import _ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m():
    cfg, w, model = make_model("qwen3", "tiny", max_ctx=64)
    yield model
    model.close()


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (128, 128, 64), (200, 256, 128), (1, 128, 64), (333, 384, 512), (129, 160, 1024), (64, 96, 2048),
                                   (2554, 2048, 2048), (700, 1536, 256)])
def test_pair_gemm_store(m, M, N, K):
    rng = np.random.default_rng(M * 7 + N + K)
    x = rng.standard_normal((M, K)).astype(np.float32)
    w = (0.05 * rng.standard_normal((N, K))).astype(np.float16)
    y, _ = m.debug_gemm(x, w, impl=4)
    ref = _ref(x, w, None, None, 0, 0)
    assert np.abs(y - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
    y2, _ = m.debug_gemm(x, w, impl=2)
    assert np.array_equal(y, y2)


@pytest.mark.parametrize("epi,act", [(1, 0), (2, 1), (2, 2), (2, 3), (3, 0)])
def test_pair_gemm_epilogues(m, epi, act):
    rng = np.random.default_rng(epi * 10 + act)
    M, N, K = 650, 768, 192
    x = rng.standard_normal((M, K)).astype(np.float32)
    w = (0.05 * rng.standard_normal((N, K))).astype(np.float16)
    bias = rng.standard_normal(N).astype(np.float32) if epi != 3 else None
    resid = rng.standard_normal((M, N)).astype(np.float32) if epi == 1 else None
    y, _ = m.debug_gemm(x, w, bias=bias, resid=resid, impl=4, epi=epi, act=act)
    ref = _ref(x, w, bias, resid, epi, act)
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= 3e-5 * max(1.0, np.abs(ref).max())
