"""Candle op semantics used by the hot path (third-party crates, see oracle/__init__.py)."""
import math
import numpy as np

try:  # scipy is in the image; keep a pure-python fallback so the oracle never silently changes
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf, otypes=[np.float32])

F32 = np.float32


def linear(x, w, b=None):
    """candle_nn::Linear::forward: x @ W^T (+ b).  w: (out, in)."""
    x = x.astype(F32, copy=False)
    w = w.astype(F32, copy=False)
    rows = int(np.prod(x.shape[:-1])) if x.ndim > 1 else 1
    if rows <= 8 and w.flags["C_CONTIGUOUS"]:
        # decode (M ~ 1): W @ x^T is the same dot products without materialising W^T (a GEMV over contiguous W)
        y = np.matmul(w, x.reshape(rows, -1).T).T.reshape(x.shape[:-1] + (w.shape[0],))
    else:
        y = np.matmul(x, w.T)
    if b is not None:
        y = y + b.astype(F32, copy=False)
    return y.astype(F32, copy=False)


def rms_norm(x, w, eps):
    """candle_nn::RmsNorm (ops::rms_norm): f32 mean of squares over the last dim."""
    x = x.astype(F32, copy=False)
    ms = np.mean(x * x, axis=-1, keepdims=True, dtype=F32)
    return (x / np.sqrt(ms + F32(eps))) * w.astype(F32, copy=False)


def layer_norm(x, w, b, eps):
    """candle_nn::LayerNorm (remove_mean=true, affine) -- modules.rs:867-875 get_layer_norm."""
    x = x.astype(F32, copy=False)
    mu = np.mean(x, axis=-1, keepdims=True, dtype=F32)
    xc = x - mu
    var = np.mean(xc * xc, axis=-1, keepdims=True, dtype=F32)
    y = xc / np.sqrt(var + F32(eps))
    y = y * w.astype(F32, copy=False)
    if b is not None:
        y = y + b.astype(F32, copy=False)
    return y.astype(F32, copy=False)


def softmax_last_dim(x):
    x = x.astype(F32, copy=False)
    m = np.max(x, axis=-1, keepdims=True)
    e = np.exp(x - m)
    return (e / np.sum(e, axis=-1, keepdims=True, dtype=F32)).astype(F32, copy=False)


def silu(x):
    x = x.astype(F32, copy=False)
    return (x / (F32(1.0) + np.exp(-x))).astype(F32, copy=False)


def gelu_erf(x):
    """Activation::Gelu -> Tensor::gelu_erf."""
    x = x.astype(F32, copy=False)
    return (F32(0.5) * x * (F32(1.0) + _erf(x * F32(1.0 / math.sqrt(2.0))).astype(F32))).astype(F32)


def gelu_tanh(x):
    """Activation::GeluPytorchTanh / NewGelu / Tensor::gelu(): tanh approximation."""
    x = x.astype(F32, copy=False)
    c = F32(math.sqrt(2.0 / math.pi))
    return (F32(0.5) * x * (F32(1.0) + np.tanh(c * (x + F32(0.044715) * x * x * x)))).astype(F32)


ACT = {
    "silu": silu,
    "swish": silu,
    "gelu": gelu_erf,
    "gelu_pytorch_tanh": gelu_tanh,
    "gelu_new": gelu_tanh,
}


def activation(name):
    return ACT[name]


def embedding(ids, table):
    return table[np.asarray(ids, dtype=np.int64)].astype(F32)


def conv2d(x, w, b, stride, padding):
    """candle_nn::Conv2d, NCHW, square stride/padding, dilation 1, groups 1 (modules.rs:815-839)."""
    n, c, h, wd = x.shape
    oc, ic, kh, kw = w.shape
    assert ic == c
    xp = np.pad(x.astype(F32, copy=False), ((0, 0), (0, 0), (padding, padding), (padding, padding)))
    oh = (h + 2 * padding - kh) // stride + 1
    ow = (wd + 2 * padding - kw) // stride + 1
    cols = np.empty((n, c, kh, kw, oh, ow), dtype=F32)
    for i in range(kh):
        for j in range(kw):
            cols[:, :, i, j] = xp[:, :, i:i + stride * oh:stride, j:j + stride * ow:stride]
    cols = cols.reshape(n, c * kh * kw, oh * ow)
    y = np.matmul(w.reshape(oc, -1).astype(F32), cols)  # (n, oc, oh*ow)
    if b is not None:
        y = y + b.astype(F32)[None, :, None]
    return y.reshape(n, oc, oh, ow).astype(F32)
